# scratch: tuning runs of the tensor-core fp32 engine (bench.py --dtype f32), one line per setting
run() {
  env "$@" timeout -s KILL 200 python bench.py --dtype f32 --steps 3 --no-cpu-baseline > gpurun_out/bench_tc_tmp.json 2> gpurun_out/bench_tc_tmp.err; rc=$?
  python - "$rc" "$*" <<'PY'
import json, sys
try:
    d = json.load(open("gpurun_out/bench_tc_tmp.json"))
    print("%-60s rc=%s value %.4g kernel_ms %.2f e2e %.4g" % (sys.argv[2], sys.argv[1], d["value"], d["roofline"]["kernel_ms"], d["e2e"]["value"]))
except Exception as e:
    print(sys.argv[2], "rc=", sys.argv[1], "parse fail", e)
PY
}
timeout -s KILL 300 python tools/gpu_check.py 2>&1 | grep -E "f32|FAILED|Error" | cut -c1-250 | sed 's/set_state.*| mean/| mean/' | grep -E "c3|c4|c5|ragged|edge|FAILED|Error"
run X=1
BOPY_B200_TC_PROF=1 timeout -s KILL 200 python bench.py --dtype f32 --steps 1 --no-cpu-baseline 2>&1 | grep tc_prof | tail -1
run BOPY_B200_TC_STAGES=2 BOPY_B200_TC_DSTAGES=4
run BOPY_B200_TC_STAGES=3 BOPY_B200_TC_DSTAGES=3
