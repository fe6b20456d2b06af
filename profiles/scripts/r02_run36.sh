for v in NOLOAD NOLOADDKGEB_EXP_NOBAR; do
KGEB_LIB_PATH=$PWD/kge-keras_b200/build_exp/libkgeb200_$v.so python profiles/exp_rank_tile.py 2> gpurun_out/r2_exp36.err | python -c "
import sys,json; d=json.loads(sys.stdin.read()); print('$v', {k:(v['tail']['kernel_ms'], v['head']['kernel_ms']) for k,v in d.items() if isinstance(v,dict)})"
done; tail -3 gpurun_out/r2_exp36.err
