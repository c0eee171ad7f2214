mkdir -p gpurun_out
TAG=$1
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/${TAG}_pytest.log
tail -4 gpurun_out/${TAG}_pytest.log
timeout 300 python bench.py --steps 200 --warmup 5 --no-cpu-baseline > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench exit $?"
python - <<P
import json
d=json.load(open('gpurun_out/${TAG}_bench.json'))
print('bench', d['ms_per_step'], d['value'], 'e2e', d['e2e']['value'], d['e2e']['ms_per_step'], 'store', d['e2e_store']['value'], d['roofline']['breakdown_ms_per_step'])
P
