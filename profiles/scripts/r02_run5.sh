set -x
python profiles/tc_numerics.py gen gpurun_out/tcnum
for e in gpurun_out/tcnum/*.in; do timeout 60 profiles/micro/tc_numerics $e ${e%.in}.out || echo "FAILED $e"; done
python profiles/tc_numerics.py analyze gpurun_out/tcnum > gpurun_out/tcnum_digest.txt 2>&1
python -m pytest tests -m gpu -x -q -k "not FB15k" > gpurun_out/r2_pytest5.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest5.log
tail -5 gpurun_out/r2_pytest5.log; tail -30 gpurun_out/tcnum_digest.txt
