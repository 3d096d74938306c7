#!/bin/bash
# First GPU run of the §8f rows written at the end of round 1 (backward kernels, frame sink, plane store staging):
#   gpurun --timeout 900 -- 'bash scripts/gpu_next_rows.sh'
# --runxfail turns the xfail(strict=False) marks off, so a failure shows its traceback; compute-sanitizer then checks
# the never-run kernels for out-of-bounds accesses; the train-step timing comes last.
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_zz_next_rows.py -q -m gpu --runxfail -x 2>&1 | tail -40 > gpurun_out/next_rows_tests.log
timeout 600 compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_zz_next_rows.py -q -m gpu --runxfail \
  -k "composite_bwd or gather_bwd_matches or frame_sink" 2>&1 | tail -30 > gpurun_out/next_rows_memcheck.log
timeout 600 python scripts/bench_train_step.py --steps 20 > gpurun_out/train_step.json 2> gpurun_out/train_step.err
timeout 600 python scripts/render_video.py --out /tmp/nvsr_video --scenes 2 --frames 8 --res 400 > gpurun_out/render_video.json 2> gpurun_out/render_video.err
timeout 600 python scripts/bench_torch_frame.py --frames 2 > gpurun_out/torch_frame.json 2> gpurun_out/torch_frame.err; cat gpurun_out/torch_frame.json
(nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I neural-volume-super-resolution_b200/csrc -I include -o /tmp/umma_wgrad scripts/ubench/umma_wgrad.cu && for k in 48 128 144; do timeout 60 /tmp/umma_wgrad $k 64 0; done; timeout 60 /tmp/umma_wgrad 128 64 1) > gpurun_out/umma_wgrad.log 2>&1
cat gpurun_out/umma_wgrad.log; tail -5 gpurun_out/next_rows_tests.log; cat gpurun_out/train_step.json gpurun_out/render_video.json
