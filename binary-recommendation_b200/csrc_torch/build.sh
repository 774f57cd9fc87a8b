#!/bin/bash
# Builds libbrk_torch.so (torch custom ops over the C ABI) in-tree, next to libbrk_b200.so.  Plain g++ with the include /
# library paths torch reports; links libbrk_b200.so through $ORIGIN so that the pair travels together.
set -e
cd "$(dirname "$0")"
PY=${PYTHON:-python}
out=../libbrk_torch.so
if [ -f "$out" ] && [ "$out" -nt brk_torch.cpp ] && [ "$out" -nt ../../include/brk_b200.h ] && [ "$out" -nt build.sh ]; then
  echo "up to date $(cd ..; pwd)/libbrk_torch.so"; exit 0
fi
INCS=$($PY -c "import torch.utils.cpp_extension as E; print(' '.join('-I' + p for p in E.include_paths()))")
LIBDIR=$($PY -c "import torch.utils.cpp_extension as E; print(E.library_paths()[0])")
ABI=$($PY -c "import torch; print(int(torch._C._GLIBCXX_USE_CXX11_ABI))")
g++ -O2 -std=c++17 -fPIC -shared -D_GLIBCXX_USE_CXX11_ABI=$ABI $INCS -I/usr/local/cuda/include brk_torch.cpp -o $out \
    -L$LIBDIR -ltorch -ltorch_cpu -lc10 -ltorch_cuda -lc10_cuda -L.. -lbrk_b200 -Wl,-rpath,'$ORIGIN' -Wl,-rpath,$LIBDIR
echo "built $(cd ..; pwd)/libbrk_torch.so"
