# Round-end check on the GPU box: tests, smoke, both bench arms, side benches and the ncu launch list of the bench command.
set -x
T=${1:-r2h}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/${T}_pytest.log 2>&1; tail -3 gpurun_out/${T}_pytest.log
timeout 300 python __graft_entry__.py > gpurun_out/${T}_smoke.log 2>&1; tail -2 gpurun_out/${T}_smoke.log
timeout 600 python bench.py > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err; tail -c 600 gpurun_out/${T}_bench.json
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${T}_bench_ref.json 2>> gpurun_out/${T}_bench.err; cat gpurun_out/${T}_bench_ref.json | tail -c 400
timeout 300 python tools/sample_bench.py binned staged > gpurun_out/${T}_sample_bench.txt 2>&1; cat gpurun_out/${T}_sample_bench.txt
timeout 300 python tools/scatter_bench.py > gpurun_out/${T}_scatter_bench.txt 2>&1; cat gpurun_out/${T}_scatter_bench.txt
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/${T}_launches.csv python bench.py --steps 3 --warmup 3 --no-cpu --no-side > gpurun_out/${T}_ncu_l.log 2>&1
ls -la gpurun_out | tail -14
