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
run BOPY_B200_TC_PREFETCH=24
run BOPY_B200_TC_PREFETCH=0
run BOPY_B200_TC_PREFETCH=24 BOPY_B200_TC_STAGES=2
run BOPY_B200_TC_PREFETCH=24 BOPY_B200_TC_FOLD=8
run BOPY_B200_TC_PREFETCH=24 BOPY_B200_TC_FOLD=2
timeout -s KILL 100 python tools/gpu_check.py c4_hart 2>&1 | grep f32 | cut -c1-220
