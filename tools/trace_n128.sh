export EO_TRACE_EXT=1
run() { echo "== $*"; env "$@" EO_TEST_STATS=1 python tools/conv3_trace.py 64 256 256 128 128 3 0
        env "$@" EO_TEST_STATS=1 python tools/conv3_trace.py 64 256 256 128 128 3 1
        env "$@" EO_TEST_STATS=1 python tools/conv3_trace.py 64 256 256 384 128 3 0
        env "$@" EO_TEST_STATS=1 python tools/conv3_trace.py 64 128 128 512 256 3 0; }
run EO_TEST_GN=2
