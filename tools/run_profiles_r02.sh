# Round-2 evidence run on one B200: bench lines (ours + reference arm), launch list and ncu captures of the
# dominant kernels.  Outputs land in gpurun_out/; summaries are copied to profiles/ by tools/collect_profiles.sh.
set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -6 > gpurun_out/tests_r02.log; tail -3 gpurun_out/tests_r02.log
python bench.py > gpurun_out/bench_r02.json 2> gpurun_out/bench_r02.err; echo rc=$?
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref_r02.json 2> gpurun_out/bench_ref_r02.err; echo rc=$?
python bench.py --steps 3 --warmup 3 --no-cpu --no-e2e --no-other > gpurun_out/plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_alt_r02.csv python bench.py --steps 3 --warmup 3 --no-cpu --no-e2e --no-other > gpurun_out/ncu_l.log 2>&1
python tools/quick_bxd.py > gpurun_out/quick5.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:scan_kernel -s 2 -c 1 -f -o gpurun_out/scan_alt_r02_final python tools/quick_bxd.py > gpurun_out/ncu_e.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:scan_kernel -s 9 -c 1 -f -o gpurun_out/scan_null_r02_final python tools/quick_bxd.py > gpurun_out/ncu_f.log 2>&1
python tools/quick_exact.py 79 7321 35554 0 4 > gpurun_out/quick_exact.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:scan_stream_kernel -s 2 -c 1 -f -o gpurun_out/scan_exact_r02 python tools/quick_exact.py 79 7321 35554 0 4 > gpurun_out/ncu_g.log 2>&1
tail -3 gpurun_out/bench_r02.err; cat gpurun_out/bench_r02.json; cat gpurun_out/bench_ref_r02.json; tail -4 gpurun_out/quick5.log
