mkdir -p gpurun_out
TAG=$1
timeout 300 python -m pytest tests/test_conv_gpu.py tests/test_fullsize_gpu.py -m gpu -x -q > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/${TAG}_pytest.log
tail -4 gpurun_out/${TAG}_pytest.log
timeout 200 python bench.py --steps 100 --warmup 5 --no-cpu-baseline > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench exit $?"
python - <<P
import json
d=json.load(open('gpurun_out/${TAG}_bench.json'))
print('bench', d['ms_per_step'], d['value'], d['e2e']['value'], d['roofline']['breakdown_ms_per_step'])
P
MOLKGNN_B200_LIB=molkgnn_b200/libmolkgnn_b200_prof.so timeout 200 python tools/phase_clocks.py 4096 10 > gpurun_out/${TAG}_phase.json 2> gpurun_out/${TAG}_phase.err; echo "phase exit $?"
python - <<P
import json
d=json.load(open('gpurun_out/${TAG}_phase.json'))
for k,v in d.items():
    if 'bwd_tile' in k or 'coef' in k: print(k, {a:round(b,1) for a,b in v.items()})
P
