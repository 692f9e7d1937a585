#!/bin/bash
# Builds the front-end with -fsanitize=address,undefined (no CUDA needed) and runs it over COUNT damaged inputs.
# usage: tools/fuzz/run_frontend_asan.sh [SEED] [COUNT]
set -e
ROOT=$(cd "$(dirname "$0")/../.." && pwd)
S=$ROOT/sdf2mesh_b200/csrc
W=$(mktemp -d)
g++ -O1 -g -std=c++17 -fsanitize=address,undefined -fno-omit-frame-pointer -I$S -I${CUDA_HOME:-/usr/local/cuda}/include \
    $ROOT/tools/fuzz/frontend_asan_driver.cpp $S/shader_api.cpp $S/frontend/*.cpp -o $W/fuzz_asan
python3 $ROOT/tools/fuzz/gen_mutations.py ${1:-1} ${2:-3000} $W/in
UBSAN_OPTIONS=print_stacktrace=1:halt_on_error=1 $W/fuzz_asan $W/in
rm -rf $W
