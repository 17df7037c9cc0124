# Round-end validation on one B200 (run under gpurun): tests, smoke, full bench line, A/B of the K2 variants, profiles.
set -x
mkdir -p gpurun_out
timeout 400 python -m pytest tests -m gpu -x -q > gpurun_out/final_tests.log 2>&1; tail -3 gpurun_out/final_tests.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 400 python bench.py > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err; tail -c 400 gpurun_out/bench_final.json
for m in 7; do
  timeout 200 python bench.py --gru-mode $m --no-cpu-baseline --no-latency --no-config3 --no-small-batch > gpurun_out/bench_mode$m.json 2> gpurun_out/bench_mode$m.err
done
SKIP_TESTS=1 timeout 700 bash scripts/prof_k1k2.sh r1c > gpurun_out/prof_r1c.out 2>&1
python - <<'PY'
import json
for f in ('bench_final', 'bench_mode7'):
    try:
        d = json.load(open('gpurun_out/%s.json' % f))
        print(f, round(d['value'] / 1e6, 1), 'M/s  K1 %.1f us  K2 %.1f us  proj %s  e2e %.1f M/s  det %d' % (
            1e3 * d['roofline']['ms_per_launch'], 1e3 * d['roofline_gru']['ms_per_launch'],
            d['roofline_gru']['input_projection_ms_per_launch'], d['e2e']['value'] / 1e6, d['detections']))
    except Exception as e:
        print(f, 'failed', e)
PY
