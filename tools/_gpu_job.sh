( echo "compute-sanitizer --tool memcheck over the WHOLE -m gpu suite (final round-2 library)";
  timeout 1500 compute-sanitizer --tool memcheck python -m pytest tests -m gpu -q -x --deselect tests/test_shared_source.py 2>&1 | grep -E "passed|failed|ERROR SUMMARY|Invalid|error" | head -20 ) > gpurun_out/r2f_sanitizer_full.txt 2>&1
cat gpurun_out/r2f_sanitizer_full.txt
