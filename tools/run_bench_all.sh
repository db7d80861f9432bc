# every bench workload on one GPU; JSON lines land in gpurun_out/bench_<workload>.json
for w in alt-grid null-grid null-exact perms scaled-null-exact; do
  python bench.py --workload $w > gpurun_out/bench_$w.json 2> gpurun_out/bench_$w.err; echo "$w rc=$?"
  tail -2 gpurun_out/bench_$w.err
done
python - <<'PY'
import json
for w in ("alt-grid","null-grid","null-exact","perms","scaled-null-exact"):
    try:
        d=json.loads(open(f"gpurun_out/bench_{w}.json").read().strip().splitlines()[-1])
        print(w, "ms/step %.3f"%d["ms_per_step"], "tests/s %.3e"%d["value"], "frac %.3f"%d["roofline"]["frac"], "kernel_ms %.3f"%d["roofline"]["kernel_ms"],
              "e2e", d["e2e"] and ("%.3e"%d["e2e"]["value"] if d["e2e"].get("value") else d["e2e"]), "cpu", d["cpu_baseline"] and "%.3e"%d["cpu_baseline"]["value"], "clk", d["clocks"])
    except Exception as e:
        print(w, "ERR", e)
PY
