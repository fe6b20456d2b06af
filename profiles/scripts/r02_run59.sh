for W in 256 64 512 128; do
echo "W=$W"
W=$W KGEB_LIB_PATH=$PWD/kge-keras_b200/build_exp/libkgeb200_oldfft.so python profiles/exp_hole_fft.py 2>/dev/null | grep "^fft" | sed 's/^/old /'
W=$W python profiles/exp_hole_fft.py 2>/dev/null | grep "^fft" | sed 's/^/new /'
done
KGEB_LIB_PATH=$PWD/kge-keras_b200/build_exp/libkgeb200_oldfft.so python profiles/exp_hole_fused.py 2>/dev/null | tail -2 | sed 's/^/old /'
python profiles/exp_hole_fused.py 2>/dev/null | tail -2 | sed 's/^/new /'
timeout 600 python -m pytest tests/test_gpu_hole_fft.py -x -q 2>&1 | tail -2
