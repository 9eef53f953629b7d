# Training-path check: the fused backward links of the MLP (tests) and the config-5 step with both MLP paths.
set -x
T=${1:-r2t}
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_train_decode.py tests/test_gpu_backward.py -x -q -m gpu > gpurun_out/${T}_pytest.log 2>&1; tail -5 gpurun_out/${T}_pytest.log
timeout 300 python tools/bench_train.py --train-precision fp16 > gpurun_out/${T}_train_fp16.json 2> gpurun_out/${T}_train.err; cat gpurun_out/${T}_train_fp16.json
timeout 300 python tools/bench_train.py --train-precision fp32 > gpurun_out/${T}_train_fp32.json 2>> gpurun_out/${T}_train.err; cat gpurun_out/${T}_train_fp32.json
TRAIN_PROFILE=1 timeout 300 python tools/bench_train.py --train-precision fp16 --steps 3 > /dev/null 2> gpurun_out/${T}_train_prof.txt; head -45 gpurun_out/${T}_train_prof.txt
