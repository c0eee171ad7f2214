TAG=$1
O=gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > $O/${TAG}_pytest.log 2>&1; echo "pytest exit $?" >> $O/${TAG}_pytest.log
tail -6 $O/${TAG}_pytest.log
timeout 300 python bench.py --steps 200 --warmup 5 --no-cpu-baseline > $O/${TAG}_bench.json 2> $O/${TAG}_bench.err; echo "bench exit $?"
python - <<P
import json
d=json.load(open('$O/${TAG}_bench.json'))
print('bench', d['ms_per_step'], d['value'], 'e2e', d['e2e']['value'], d['e2e']['ms_per_step'], 'store', d['e2e_store']['value'])
P
python tools/host_profile.py host > $O/${TAG}_host.txt 2>&1; grep -v "^$" $O/${TAG}_host.txt | head -24 | cut -c1-150
MOLKGNN_B200_LIB=molkgnn_b200/libmolkgnn_b200_prof.so timeout 200 python tools/phase_clocks.py 4096 10 > $O/${TAG}_phase.json 2> $O/${TAG}_phase.err
python -c "
import json
d=json.load(open('$O/${TAG}_phase.json'))
print({k:v for k,v in d.items() if not isinstance(v,dict)})"
