import json,sys
for f in sys.argv[1:]:
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split('/')[-1], "value=%.3e"%d["value"], "e2e=%.3e"%d["e2e"]["value"], {k:round(v,4) for k,v in d["kernel_ms"].items()}, "eval=%.3e"%d["eval"]["value"])
    except Exception as e: print(f, "ERR", e, open(f).read()[-800:])
