for rep in 1 2; do
KGEB_LIB_PATH=$PWD/kge-keras_b200/build_exp/libkgeb200_oldring.so python profiles/exp_train.py 2>/dev/null | grep "margin 1.0" | sed 's/^/old /'
python profiles/exp_train.py 2>/dev/null | grep "margin 1.0" | sed 's/^/new /'
done
timeout 600 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_dropout.py tests/test_gpu_fullsize.py tests/test_gpu_sortseg.py -x -q -k "train or step or dropout or fullsize or sortseg" 2>&1 | tail -2
