# usage: bash scripts/regs_probe.sh [extra nvcc flags]  -> registers / spills of the C2 instantiation only (10 s)
root=$(cd "$(dirname "$0")/.." && pwd)
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 --extended-lambda -DMCL_ONLY_C2 "$@" \
     -I $root/include -I $root/mcluminescence_b200/csrc -Xptxas -v -c $root/mcluminescence_b200/csrc/mcl_philox.cu -o /tmp/probe.o 2>&1 | grep -A2 "philox_kernelILi256" | grep -E "spill|registers" | tr '\n' ' '; echo
