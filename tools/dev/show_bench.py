import json,sys
d=json.loads(sys.stdin.read()); print({k:round(d[k],4) for k in ("value","ms_per_step","ms_per_step_median","value_single_stream")}, round(d["e2e"]["value"]), {k:round(v["ms"],4) for k,v in d["kernels"].items()}, d["clocks"]["sm_mhz"])
