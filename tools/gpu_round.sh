#!/bin/bash
# One GPU session: parity tests, bench lines of the configs, ncu launch list and a full capture of the step's main kernels,
# in-kernel phase clocks.  Usage (under gpurun): bash tools/gpu_round.sh <tag>
TAG=${1:-r02Z}
O=gpurun_out
mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $O/${TAG}_smi.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > $O/${TAG}_pytest.log 2>&1; echo "pytest exit $?" >> $O/${TAG}_pytest.log
tail -4 $O/${TAG}_pytest.log
timeout 600 python bench.py --steps 200 --warmup 5 > $O/${TAG}_bench.json 2> $O/${TAG}_bench.err; echo "bench exit $?"
timeout 600 python bench.py --impl reference --steps 4 --warmup 1 > $O/${TAG}_bench_reference.json 2> $O/${TAG}_bench_reference.err; echo "reference exit $?"
timeout 300 python bench.py --molecules 65536 --steps 20 --warmup 3 --no-cpu-baseline > $O/${TAG}_bench_65536.json 2> $O/${TAG}_bench_65536.err; echo "65536 exit $?"
timeout 300 python bench.py --molecules 65536 --forward-only --steps 30 --warmup 3 --no-cpu-baseline > $O/${TAG}_bench_fwdonly_65536.json 2> $O/${TAG}_bench_fwdonly.err; echo "fwdonly exit $?"
timeout 300 python bench.py --wide --steps 20 --warmup 3 --no-cpu-baseline > $O/${TAG}_bench_wide.json 2> $O/${TAG}_bench_wide.err; echo "wide exit $?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/${TAG}_launches.csv \
    python tools/prof_step.py 4096 3 > $O/${TAG}_ncu_launches.log 2>&1; echo "ncu launches exit $?"
timeout 900 ncu --set full --clock-control none --import-source on \
    -k regex:'k_conv_bwd_tile|k_coef_tile|k_stack_fwd_fused|k_param_finalize' -s 20 -c 10 \
    -f -o $O/${TAG}_prof python tools/prof_step.py 4096 3 > $O/${TAG}_ncu_full.log 2>&1; echo "ncu full exit $?"
MOLKGNN_B200_LIB=molkgnn_b200/libmolkgnn_b200_prof.so timeout 200 python tools/phase_clocks.py 4096 10 > $O/${TAG}_phase.json 2> $O/${TAG}_phase.err; echo "phase exit $?"
python - <<P
import json
for n in ("bench", "bench_65536", "bench_fwdonly_65536", "bench_wide"):
    try:
        d = json.load(open("$O/${TAG}_%s.json" % n))
        print(n, round(d["ms_per_step"], 4), round(d["value"]), "e2e", round(d["e2e"]["value"]), d["roofline"]["kernel"], round(d["roofline"]["frac"], 4), round(d["roofline"]["step_frac"], 4), d.get("gpu_launches"))
    except Exception as e:
        print(n, "failed", e)
P
ls -la $O | tail -5
