#!/bin/bash
# Host code of libpsb200.so under AddressSanitizer + UBSan, no GPU needed: newton.cpp, market.cpp and the host parts of
# dist.cu (plan builder), solver.cu (parameters, validation) and capi.cu are rebuilt instrumented into a scratch directory, linked with the regular device
# objects, and the CPU tests that exercise them (Newton driver vs the restatement, Matrix Market reader incl. hostile
# files, host plan incl. corrupted index arrays, parameter parser) run against that library.
set -e
cd "$(dirname "$0")/.."
OUT=gpurun_out/asan
mkdir -p $OUT
C=polysolve_b200/csrc
make -C $C libpsb200.so > /dev/null
SAN="-fsanitize=address,undefined -fno-omit-frame-pointer"
for f in newton market; do /usr/bin/g++ -O1 -g -std=c++17 -fPIC $SAN -c $C/$f.cpp -o $OUT/$f.o; done
for f in dist capi solver; do
  /usr/local/cuda/bin/nvcc -O1 -std=c++17 -gencode arch=compute_100a,code=sm_100a --expt-relaxed-constexpr -ccbin /usr/bin/g++ \
    -Xcompiler -fPIC,-fsanitize=address,-fno-omit-frame-pointer -c $C/$f.cu -o $OUT/$f.o
done
/usr/local/cuda/bin/nvcc -shared -gencode arch=compute_100a,code=sm_100a -ccbin /usr/bin/g++ -Xcompiler -fsanitize=address -o $OUT/libpsb200.so \
  $OUT/capi.o $OUT/solver.o $C/amg.o $C/amg_dist.o $C/spgemm.o $C/dense.o $OUT/dist.o $C/fem.o $C/lbfgs.o $C/neohookean.o $OUT/newton.o $OUT/market.o \
  -lcudart_static -ldl -lrt -lpthread -lubsan
cat > $OUT/run.py <<'PY'
import sys
sys.path.insert(0, '.')
import polysolve_b200._lib as L
L.LIB_PATH = 'gpurun_out/asan/libpsb200.so'
import pytest
sys.exit(pytest.main(['tests/test_newton.py', 'tests/test_market_io.py', 'tests/test_dist_cpu.py', 'tests/test_capi_cpu.py', '-q', '-x', '-s',
                      '-m', 'not gpu', '-p', 'no:cacheprovider', '-k', 'not gloo and not headers_are_plain_c and not fails_loudly']))
PY
GCCLIB=$(dirname "$(gcc -print-file-name=libasan.so)")
LD_PRELOAD="$GCCLIB/libasan.so $(gcc -print-file-name=libstdc++.so.6)" ASAN_OPTIONS=detect_leaks=0:halt_on_error=1 \
  UBSAN_OPTIONS=print_stacktrace=1:halt_on_error=1 python $OUT/run.py 2>&1 | tee $OUT/out.log | tail -3
# the CPU oracle (test infrastructure) under the same sanitizers: a memory error there would undermine every parity claim
/usr/bin/g++ -O1 -g -march=x86-64-v3 -std=c++17 -fopenmp -fPIC $SAN -shared -o $OUT/liboracle.so oracle/oracle.cpp oracle/amg_oracle.cpp
cat > $OUT/run_oracle.py <<'PY'
import os
import sys
sys.path.insert(0, '.')
import oracle.oracle as O
O._HERE = os.path.abspath('gpurun_out/asan')      # load the instrumented liboracle.so
O.build = lambda: None
import pytest
sys.exit(pytest.main(['tests/test_oracle.py', 'tests/test_dist_amg_cpu.py', 'tests/test_fem.py', 'tests/test_saddle_point.py', '-q', '-x', '-s',
                      '-m', 'not gpu', '-p', 'no:cacheprovider', '-k', 'not gloo']))
PY
LD_PRELOAD="$GCCLIB/libasan.so $(gcc -print-file-name=libstdc++.so.6)" ASAN_OPTIONS=detect_leaks=0:halt_on_error=1 \
  UBSAN_OPTIONS=print_stacktrace=1:halt_on_error=1 OMP_NUM_THREADS=4 python $OUT/run_oracle.py 2>&1 | tee $OUT/out_oracle.log | tail -2
# ThreadSanitizer on the multi-threaded host plan builder (dist.cu: DistPlanHost::build splits its passes over host threads
# above 2^20 entries)
T=gpurun_out/tsan
mkdir -p $T
/usr/local/cuda/bin/nvcc -O1 -std=c++17 -gencode arch=compute_100a,code=sm_100a --expt-relaxed-constexpr -ccbin /usr/bin/g++ \
  -Xcompiler -fPIC,-fsanitize=thread,-g -c $C/dist.cu -o $T/dist.o
/usr/local/cuda/bin/nvcc -shared -gencode arch=compute_100a,code=sm_100a -ccbin /usr/bin/g++ -Xcompiler -fsanitize=thread -o $T/libpsb200.so \
  $C/capi.o $C/solver.o $C/amg.o $C/amg_dist.o $C/spgemm.o $C/dense.o $T/dist.o $C/fem.o $C/lbfgs.o $C/neohookean.o $C/newton.o $C/market.o \
  -lcudart_static -ldl -lrt -lpthread
cat > $T/run.py <<'PY'
import sys
sys.path.insert(0, '.')
import polysolve_b200._lib as L
L.LIB_PATH = 'gpurun_out/tsan/libpsb200.so'
import pytest
sys.exit(pytest.main(['tests/test_dist_cpu.py', '-q', '-x', '-s', '-m', 'not gpu', '-p', 'no:cacheprovider', '-k', 'threaded or randomised or bit_exact']))
PY
LD_PRELOAD="$(gcc -print-file-name=libtsan.so) $(gcc -print-file-name=libstdc++.so.6)" TSAN_OPTIONS="halt_on_error=0 report_signal_unsafe=0" \
  python $T/run.py 2>&1 | tee $T/out.log | tail -2
if grep -q "ERROR: AddressSanitizer\|runtime error" $OUT/out.log $OUT/out_oracle.log || grep -q "WARNING: ThreadSanitizer" $T/out.log; then echo "SANITIZER FINDINGS"; exit 1; fi
echo "sanitizers: clean"
