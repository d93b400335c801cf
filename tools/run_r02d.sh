#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_tc.py tests/test_gpu_unet.py -q -p no:cacheprovider -x 2>&1 | tail -3
EO_B200_LIB=$PWD/eo_diffusion_b200/libeo_b200_dev.so EO_TEST_GN=2 EO_TEST_STATS=1 EO_TRACE_EXT=1 timeout 300 python tools/conv3_trace.py > gpurun_out/conv3_trace_r02d.log 2>&1
cat gpurun_out/conv3_trace_r02d.log
bash tools/ab.sh r02d prev base
