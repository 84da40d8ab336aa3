#!/bin/bash
# One gpurun call that confirms a round's final state on a B200 and takes the profile evidence for it:
#   the GPU suite, smoke(), the default bench line, and ncu --set full captures of replace_stream_kernel<false>
#   (the first passes that send their pair-count deltas to global memory, config 3; a mid-run batched pass, config 2).
# Everything lands in gpurun_out/<tag>_*.  usage: gpurun --timeout 600 -- 'bash tools/gpu_round_check.sh r2f'
tag=${1:-check}
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q --durations=12 > gpurun_out/${tag}_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/${tag}_pytest.log
tail -3 gpurun_out/${tag}_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${tag}_smoke.log 2>&1; echo "rc=$?" >> gpurun_out/${tag}_smoke.log
tail -2 gpurun_out/${tag}_smoke.log
python bench.py > gpurun_out/${tag}_bench_n1.json 2> gpurun_out/${tag}_bench_n1.err; echo "rc=$?" >> gpurun_out/${tag}_bench_n1.err
head -c 600 gpurun_out/${tag}_bench_n1.json; echo
# (profiles only after the same commands have run without ncu: the suite and the bench above cover both workloads)
timeout 150 ncu --set full --clock-control none --import-source on --kernel-name-base mangled -k regex:replace_stream_kernelILb0 -s 0 -c 3 \
    -o gpurun_out/${tag}_c3_first_global_delta_passes python tools/phase_timers.py c3 400 --plain > gpurun_out/${tag}_ncu_c3.log 2>&1
timeout 150 ncu --set full --clock-control none --import-source on --kernel-name-base mangled -k regex:replace_stream_kernelILb0 -s 300 -c 1 \
    -o gpurun_out/${tag}_c2_batched_pass python tools/phase_timers.py c2 --plain > gpurun_out/${tag}_ncu_c2.log 2>&1
ls -la gpurun_out | tail -12
