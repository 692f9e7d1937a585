#!/bin/bash
# usage: tools/fuzz/run_writers_sanitizers.sh   -- TSan + ASan/UBSan over the writer pipeline; prints one checksum line per configuration
set -e
ROOT=$(cd "$(dirname "$0")/../.." && pwd)
S=$ROOT/sdf2mesh_b200/csrc
W=$(mktemp -d)
cd $W
g++ -O1 -g -std=c++17 -fsanitize=thread -I$S -I${CUDA_HOME:-/usr/local/cuda}/include $ROOT/tools/fuzz/writers_sanitizer_driver.cpp $S/writers.cpp -o w_tsan -lpthread
g++ -O1 -g -std=c++17 -fsanitize=address,undefined -I$S -I${CUDA_HOME:-/usr/local/cuda}/include $ROOT/tools/fuzz/writers_sanitizer_driver.cpp $S/writers.cpp -o w_asan -lpthread
for cfg in "8 7" "3 100" "16 1000" "2 50000"; do
  set -- $cfg
  S2M_WRITER_THREADS=$1 S2M_WRITER_CHUNK=$2 ./w_tsan | tail -1
  echo "threads $1 chunk $2: $(md5sum a.stl a.ply b.stl | cut -c1-12 | tr '\n' ' ')"
  S2M_WRITER_THREADS=$1 S2M_WRITER_CHUNK=$2 ./w_asan | tail -1
done
cd /; rm -rf $W
