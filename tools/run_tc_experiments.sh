# scratch: quick A/B of the sweep kernels (bench.py --headline-only), one line per setting
run() {
  env "$@" timeout -s KILL 300 python bench.py --headline-only --steps 3 --no-cpu-baseline $EXTRA > gpurun_out/bench_tc_tmp.json 2> gpurun_out/bench_tc_tmp.err; rc=$?
  python - "$rc" "$* $EXTRA" <<'PY'
import json, sys
try:
    d = json.load(open("gpurun_out/bench_tc_tmp.json"))
    print("%-60s rc=%s value %.4g kernel_ms %.2f e2e %.4g frac %.3f" % (sys.argv[2], sys.argv[1], d["value"], d["roofline"]["kernel_ms"], d["e2e"]["value"], d["roofline"]["frac"]))
except Exception as e:
    print(sys.argv[2], "rc=", sys.argv[1], "parse fail", e)
PY
}
timeout -s KILL 300 python tools/gpu_check.py c4_ 2>&1 | grep -E "f32 |f64 |FAILED|Error" | cut -c1-250 | sed 's/set_state.*| mean/| mean/'
EXTRA="" run X=1
EXTRA="--dtype f32" run X=1
BOPY_B200_TC_PROF=1 timeout -s KILL 200 python bench.py --dtype f32 --steps 1 --no-cpu-baseline --headline-only 2>&1 | grep tc_prof | tail -1
EXTRA="--n 256 --d 2 --candidates 1048576" run X=1
EXTRA="--n 8192 --d 20 --candidates 262144" run X=1
