for nt in 64 128 256; do echo "in-tree NT=$nt"; MCL_PHILOX_NT=$nt bash scripts/ab_one.sh mcluminescence_b200/_lib/libmcl_b200.so c5 5920; done
echo "in-tree NT=64 PIPE=0"; MCL_PHILOX_PIPE=0 bash scripts/ab_one.sh mcluminescence_b200/_lib/libmcl_b200.so c5 5920
echo "p128_6 NT=128"; MCL_PHILOX_NT=128 bash scripts/ab_one.sh scripts/ab_libs/p128_6.so c5 5920
echo "p128_8 NT=128"; MCL_PHILOX_NT=128 bash scripts/ab_one.sh scripts/ab_libs/p128_8.so c5 5920
echo "p128_8 NT=64"; MCL_PHILOX_NT=64 bash scripts/ab_one.sh scripts/ab_libs/p128_8.so c5 5920
echo "c2 in-tree"; bash scripts/ab_one.sh mcluminescence_b200/_lib/libmcl_b200.so c2 2960
