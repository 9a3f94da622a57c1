set +e
S='python -c "import __graft_entry__ as g; g.smoke()"'
for t in memcheck racecheck synccheck initcheck; do
  echo "=== $t ===" >> gpurun_out/sanitizer.log
  timeout 200 compute-sanitizer --tool $t --print-limit 20 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | grep -v "^$" | tail -25 >> gpurun_out/sanitizer.log
  echo "rc=$?" >> gpurun_out/sanitizer.log
done
echo "=== memcheck pytest error_behaviour + carry_over + remap + fast ===" >> gpurun_out/sanitizer.log
timeout 420 compute-sanitizer --tool memcheck --print-limit 20 python -m pytest -m gpu -x -q tests/test_error_behaviour.py tests/test_carry_over.py tests/test_remap.py tests/test_fast_detect.py tests/test_geometry_validation.py 2>&1 | tail -25 >> gpurun_out/sanitizer.log
cat gpurun_out/sanitizer.log | tail -60
