# usage: bash scripts/build_variant.sh <name> [nvcc flags for mcl_philox.cu, e.g. -DMCL_ONE_CHAINS=3]
# Builds scripts/ab_libs/<name>.so = the in-tree library with mcl_philox.cu recompiled under the given flags
# (A/B runs: MCL_B200_LIB=$PWD/scripts/ab_libs/<name>.so python bench.py ...).  Needs a prior `python -m mcluminescence_b200.build`.
set -e
name=$1; shift
root=$(cd "$(dirname "$0")/.." && pwd)
lib=$root/mcluminescence_b200/_lib; src=$root/mcluminescence_b200/csrc
mkdir -p $root/scripts/ab_libs
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 --extended-lambda -Xcompiler -fPIC \
     -I $root/include -I $src "$@" -Xptxas -v -c $src/mcl_philox.cu -o /tmp/mcl_philox_$name.o 2> /tmp/mcl_philox_$name.log
grep -A2 "philox_kernelILi256ELi[0-9]*EtLi2ELb0ELb0E\|philox_kernelILi64ELi[0-9]*EtLi2ELb0ELb0E" /tmp/mcl_philox_$name.log | grep -E "spill|registers" | tr '\n' ' '; echo
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o $root/scripts/ab_libs/$name.so \
     $lib/mcl_abi.o $lib/mcl_replay.o /tmp/mcl_philox_$name.o $lib/mcl_smallbox.o $lib/mcl_peaks.o $lib/mcl_objective.o
echo built scripts/ab_libs/$name.so
