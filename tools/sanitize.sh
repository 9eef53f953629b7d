# compute-sanitizer pass over a subset of the GPU parity tests (small shapes: the tool slows kernels ~10-100x).
# usage: tools/sanitize.sh <memcheck|racecheck|synccheck> [tag]
TOOL=${1:-memcheck}
T=${2:-r2}
SUB="tests/test_gpu_parity.py tests/test_gpu_edge_cases.py tests/test_gpu_sampler_binned.py"
KEXPR="not full and not cfg4 and not large"
mkdir -p gpurun_out
timeout 600 python -m pytest $SUB -q -m gpu -x -k "$KEXPR" > gpurun_out/${T}_san_plain.log 2>&1 || { tail -5 gpurun_out/${T}_san_plain.log; exit 1; }
tail -1 gpurun_out/${T}_san_plain.log
timeout 2400 compute-sanitizer --tool $TOOL --error-exitcode 86 --log-file gpurun_out/${T}_sanitizer_${TOOL}.log \
    python -m pytest $SUB -q -m gpu -x -k "$KEXPR" > gpurun_out/${T}_san_${TOOL}_pytest.log 2>&1
echo "exit $?"
tail -2 gpurun_out/${T}_san_${TOOL}_pytest.log
grep -c "=========" gpurun_out/${T}_sanitizer_${TOOL}.log; tail -5 gpurun_out/${T}_sanitizer_${TOOL}.log
