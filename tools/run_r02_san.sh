# compute-sanitizer on the C smoke program (small alt-grid scan, permutation scan, thresholds, error paths):
# memcheck, then racecheck (shared-memory hazards of the hand-rolled mbarrier ring / ping-pong turns), then synccheck.
set -x
python -c "import sys; sys.path.insert(0, 'tests'); import test_cabi_smoke as t; print(t.build())"
EXE=bulklmm.jl_b200/blmm_b200/lib/cabi_smoke
for tool in memcheck racecheck synccheck; do
  timeout 600 compute-sanitizer --tool $tool --print-limit 20 $EXE tests/golden/cabi_smoke.bin > gpurun_out/sanitizer_$tool.log 2>&1; echo "$tool rc=$?"
  tail -4 gpurun_out/sanitizer_$tool.log
done
python tools/quick_bxd.py 2>&1 | grep -E "null-grid|alt-grid:" | tail -3
python tools/quick_perms.py 2>&1 | tail -1
