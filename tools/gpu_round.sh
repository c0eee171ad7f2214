#!/bin/bash
# One GPU session: parity tests, bench line, ncu launch list and a full capture of every kernel of one step.
# Usage (under gpurun): bash tools/gpu_round.sh <tag>
TAG=${1:-r01}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/${TAG}_smi.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/${TAG}_pytest.log
tail -5 gpurun_out/${TAG}_pytest.log
timeout 600 python bench.py --steps 30 --warmup 5 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench exit $?"
cat gpurun_out/${TAG}_bench.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${TAG}_launches.csv \
    python tools/prof_step.py 4096 2 > gpurun_out/${TAG}_ncu_launches.log 2>&1; echo "ncu launches exit $?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_conv|k_bwd|k_propagate|k_coef|k_x_images' -s 16 -c 16 \
    -f -o gpurun_out/${TAG}_prof python tools/prof_step.py 4096 2 > gpurun_out/${TAG}_ncu_full.log 2>&1; echo "ncu full exit $?"
ls -la gpurun_out
