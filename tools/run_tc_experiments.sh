run() {
  env "$@" timeout -s KILL 200 python bench.py --headline-only --steps 3 --no-cpu-baseline --dtype f32 > gpurun_out/bench_tc_tmp.json 2> gpurun_out/bench_tc_tmp.err; rc=$?
  python - "$rc" "$*" <<'PY'
import json, sys
try:
    d = json.loads(open("gpurun_out/bench_tc_tmp.json").read().strip().splitlines()[-1])
    print("%-60s rc=%s value %.4g kernel_ms %.2f frac %.3f" % (sys.argv[2], sys.argv[1], d["value"], d["roofline"]["kernel_ms"], d["roofline"]["frac"]))
except Exception as e:
    print(sys.argv[2], "rc=", sys.argv[1], "parse fail", e)
PY
}
run X=1
run BOPY_B200_TC_STAGES=3 BOPY_B200_TC_DSTAGES=4
run BOPY_B200_TC_STAGES=3 BOPY_B200_TC_DSTAGES=3
run BOPY_B200_TC_STAGES=2 BOPY_B200_TC_DSTAGES=4
BOPY_B200_TC_PROF=1 timeout -s KILL 200 python bench.py --dtype f32 --steps 1 --no-cpu-baseline --headline-only 2>&1 | grep tc_prof | tail -1
