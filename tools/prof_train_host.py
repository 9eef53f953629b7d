"""Host-side profile (cProfile) of the config-5 training step: where the launch-rate-bound step spends its host time.
usage: python tools/prof_train_host.py [--adam fused]"""
import cProfile
import os
import pstats
import runpy
import sys

sys.argv = [os.path.join(os.path.dirname(os.path.abspath(__file__)), "bench_train.py"), "--steps", "30"] + sys.argv[1:]
pr = cProfile.Profile()
pr.enable()
try:
    runpy.run_path(sys.argv[0], run_name="__main__")
finally:
    pr.disable()
    st = pstats.Stats(pr, stream=sys.stderr)
    st.sort_stats("cumulative").print_stats(70)
