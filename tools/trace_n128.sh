for gn in 0 2; do for st in 0 1; do
echo "== GN=$gn STATS=$st"
EO_TEST_GN=$gn EO_TEST_STATS=$st python tools/conv3_trace.py 64 256 256 128 128 3 0
EO_TEST_GN=$gn EO_TEST_STATS=$st python tools/conv3_trace.py 64 256 256 128 128 3 1
EO_TEST_GN=$gn EO_TEST_STATS=$st python tools/conv3_trace.py 64 256 256 384 128 3 0
EO_TEST_GN=$gn EO_TEST_STATS=$st python tools/conv3_trace.py 64 128 128 256 256 3 0
done; done
