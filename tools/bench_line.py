import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(round(d["ms_per_step"],3), d["state_hash"], {k:round(v,3) for k,v in d["phase_ms"].items()})
