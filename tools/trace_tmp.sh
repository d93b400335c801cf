export EO_TRACE_EXT=1
EO_TEST_GN=2 EO_TEST_STATS=1 python tools/conv3_trace.py 64 256 256 128 128 3 0
EO_TEST_GN=2 EO_TEST_STATS=1 python tools/conv3_trace.py 64 256 256 384 128 3 0
unset EO_TRACE_EXT
python -m pytest tests/test_gpu_tc.py tests/test_gpu_unet.py -x -q 2>&1 | tail -1
python bench.py --steps 8 --warmup 3 --no-cpu 2>&1 | grep -E "k_conv_tc3|k_attn|ms_per_step" | cut -c1-250
