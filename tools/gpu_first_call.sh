#!/bin/bash
# First GPU call of a round:  gpurun --timeout 1500 -- 'bash tools/gpu_first_call.sh'
# 1. the tests that have never run on a device (reference-graph golden vectors, sampled KL),
# 2. the whole GPU suite, 3. one bench line, 4. the launch list of a short bench run.
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1
timeout 600 python -m pytest tests/test_zz_gpu_reference_graph.py -m gpu -q 2>&1 | tail -40 > gpurun_out/zz_tests.log
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/gpu_tests.log 2>&1
tail -5 gpurun_out/gpu_tests.log
timeout 600 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err
cat gpurun_out/bench.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
    --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline \
    > gpurun_out/bench_under_ncu.log 2>&1
cat gpurun_out/zz_tests.log
