"""Diagnostic: run descriptor variants of the tcgen05 self test and print the error of each.
   python tools/tc_probe.py a_mn b_mn swap"""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tests.test_tc_gpu import run
a_mn, b_mn, swap = (int(v) for v in sys.argv[1:4])
for N, K in ((16, 16), (64, 64), (112, 112), (208, 128)):
    try:
        err, mag = run(N, K, a_mn, b_mn, swap)
    except Exception as e:
        err, mag = str(e)[:200], 0
    print(json.dumps(dict(a_mn=a_mn, b_mn=b_mn, swap=swap, N=N, K=K, err=err, mag=mag)), flush=True)
