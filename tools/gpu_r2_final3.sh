# Last check of round 2 on one GPU: all GPU tests, smoke, the bench line, the config-5 step.
set -x
T=${1:-r2w}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/${T}_pytest.log 2>&1; tail -3 gpurun_out/${T}_pytest.log
timeout 300 python __graft_entry__.py > gpurun_out/${T}_smoke.log 2>&1; tail -2 gpurun_out/${T}_smoke.log
timeout 600 python bench.py --no-side > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err; tail -c 200 gpurun_out/${T}_bench.json
for a in foreach fused; do timeout 300 python tools/bench_train.py --adam $a > gpurun_out/${T}_train_$a.json 2> gpurun_out/${T}_train.err; cut -c1-200 gpurun_out/${T}_train_$a.json; done
