# r2 first check: GPU tests, smoke, bench (N=1), reference arm
set -x
T=r2a
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/${T}_pytest.log 2>&1; tail -15 gpurun_out/${T}_pytest.log
timeout 300 python __graft_entry__.py > gpurun_out/${T}_smoke.log 2>&1; tail -2 gpurun_out/${T}_smoke.log
timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err; tail -c 3000 gpurun_out/${T}_bench.json; tail -5 gpurun_out/${T}_bench.err
timeout 600 python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/${T}_bench_ref.json 2>> gpurun_out/${T}_bench.err; tail -c 600 gpurun_out/${T}_bench_ref.json
