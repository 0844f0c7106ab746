#!/bin/bash
# One GPU-box pass: full GPU test suite, the default bench line, ncu shape metrics of the list kernel, cold per-shape timings.
#   /usr/local/graft/bin/gpurun --timeout 460 -- 'bash tools/round_check.sh r02b'
tag=${1:-check}
mkdir -p gpurun_out
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__inst_executed.avg.per_cycle_elapsed,smsp__inst_executed.sum,l1tex__data_pipe_lsu_wavefronts_mem_shared.sum,l1tex__data_pipe_lsu_wavefronts_mem_shared_op_ld.sum,l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_ld.sum,l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed,sm__throughput.avg.pct_of_peak_sustained_elapsed,launch__registers_per_thread
timeout 200 python -m pytest tests -m gpu -x -q > gpurun_out/${tag}_tests.txt 2>&1
tail -3 gpurun_out/${tag}_tests.txt
timeout 120 python bench.py > gpurun_out/${tag}_bench_n1.json 2> gpurun_out/${tag}_bench_n1.err
timeout 100 ncu --metrics $M --clock-control none -k regex:gemv_lists --csv --log-file gpurun_out/${tag}_ncu_lists.csv \
    python tools/profile_gemv.py --once b24_k65536_r256 > gpurun_out/${tag}_ncu.out 2>&1
timeout 60 python tools/profile_gemv.py b24_k65536_r256 > gpurun_out/${tag}_shapes.json 2> gpurun_out/${tag}_shapes.err
tail -c 700 gpurun_out/${tag}_bench_n1.json
