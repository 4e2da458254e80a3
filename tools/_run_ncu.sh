# usage: LIB=build_var/x.so TAG=name bash tools/_run_ncu.sh
export BNP_LIB=$PWD/${LIB:-plonky2_bn254_pairing_b200/libbnp.so}
export BNP_PHASE_MODE=${PM:-1}
timeout 600 ncu --set full --clock-control none --import-source on -k regex:bnp_vm_kernel -c 1 -o gpurun_out/prof_${TAG} -f python bench.py --steps 1 --warmup 0 --no-cpu > gpurun_out/ncu_${TAG}.log 2>&1
tail -2 gpurun_out/ncu_${TAG}.log
