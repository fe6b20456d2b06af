python profiles/tc_numerics.py gen gpurun_out/tcnum2 > /dev/null
for e in random_uniform_bf16_x1 random_uniform_fp16_x1 align_bf16; do
  timeout 60 profiles/micro/tc_numerics gpurun_out/tcnum2/$e.in gpurun_out/tcnum2/$e.out
  TC_A_TMEM=1 timeout 60 profiles/micro/tc_numerics gpurun_out/tcnum2/$e.in gpurun_out/tcnum2/$e.tmem.out; echo "rc=$?"
done
python - <<'PY'
import numpy as np
for e in ("random_uniform_bf16_x1","random_uniform_fp16_x1","align_bf16"):
    a=np.fromfile(f"gpurun_out/tcnum2/{e}.out",dtype=np.float32); b=np.fromfile(f"gpurun_out/tcnum2/{e}.tmem.out",dtype=np.float32)
    print(e, "identical:", np.array_equal(a,b), "max|diff|", float(np.abs(a-b).max()), "nonzero", int((b!=0).sum()), a[:4], b[:4])
PY
rm -rf gpurun_out/tcnum2
