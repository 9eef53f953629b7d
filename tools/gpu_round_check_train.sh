# Round-end check on one GPU: all GPU tests, smoke, both bench arms, the config-5 training step, its launch list and an ncu capture of the link kernel.
set -x
T=${1:-r2z}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/${T}_pytest.log 2>&1; tail -3 gpurun_out/${T}_pytest.log
timeout 300 python __graft_entry__.py > gpurun_out/${T}_smoke.log 2>&1; tail -2 gpurun_out/${T}_smoke.log
timeout 600 python bench.py > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err; tail -c 300 gpurun_out/${T}_bench.json
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${T}_bench_ref.json 2>> gpurun_out/${T}_bench.err; tail -c 200 gpurun_out/${T}_bench_ref.json
timeout 300 python tools/bench_train.py > gpurun_out/${T}_train_fp16.json 2> gpurun_out/${T}_train.err; cut -c1-200 gpurun_out/${T}_train_fp16.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/${T}_train_launches.csv python tools/bench_train.py --steps 1 --warmup 1 > gpurun_out/${T}_ncu_t.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:mlp_grad_link -s 4 -c 2 -o gpurun_out/${T}_link python tools/bench_train.py --steps 1 --warmup 1 > gpurun_out/${T}_ncu_link.log 2>&1
ls -la gpurun_out | tail -8
