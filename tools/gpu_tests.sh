# all GPU tests, no -x (see every failure)
T=${1:-r2}
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -q -m gpu > gpurun_out/${T}_pytest.log 2>&1; tail -40 gpurun_out/${T}_pytest.log
