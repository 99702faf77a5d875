# memcheck / racecheck / synccheck of the new kernels at test sizes (compute-sanitizer slows kernels down 10-100x: small selections only)
export PYTHONPATH=.
for tool in memcheck racecheck synccheck; do
  echo "== $tool: tests/test_em_tc_gpu.py (two cases)"
  timeout 600 compute-sanitizer --tool $tool --error-exitcode 9 python -m pytest tests/test_em_tc_gpu.py -x -q -k "200-13-5000 or underflowed or 65-50" 2>&1 | tail -6
done
echo "== memcheck: scorer + pipeline"
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gmm_tc_gpu.py tests/test_pipeline_gpu.py -x -q -k "not ragged and not full_size" 2>&1 | tail -6
