# Round-end check on the GPU box: all GPU tests, smoke, both bench arms, config-5 training step, side benches, launch lists.
set -x
T=${1:-r2f}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/${T}_pytest.log 2>&1; tail -3 gpurun_out/${T}_pytest.log
timeout 300 python __graft_entry__.py > gpurun_out/${T}_smoke.log 2>&1; tail -2 gpurun_out/${T}_smoke.log
timeout 600 python bench.py > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err; tail -c 700 gpurun_out/${T}_bench.json
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${T}_bench_ref.json 2>> gpurun_out/${T}_bench.err; tail -c 300 gpurun_out/${T}_bench_ref.json
for a in foreach fused; do
timeout 300 python tools/bench_train.py --train-precision fp16 --adam $a > gpurun_out/${T}_train_fp16_$a.json 2> gpurun_out/${T}_train.err; cut -c1-120 gpurun_out/${T}_train_fp16_$a.json
done
timeout 300 python tools/bench_train.py --train-precision fp32 > gpurun_out/${T}_train_fp32.json 2>> gpurun_out/${T}_train.err; cut -c1-120 gpurun_out/${T}_train_fp32.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/${T}_train_launches.csv python tools/bench_train.py --train-precision fp16 --steps 1 --warmup 1 > gpurun_out/${T}_ncu_t.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/${T}_launches.csv python bench.py --steps 3 --warmup 3 --no-cpu --no-side > gpurun_out/${T}_ncu_l.log 2>&1
ls -la gpurun_out | tail -14
