mkdir -p gpurun_out
for m in 0 1; do
SHAPE=512,64,16,16,128 SN_IGEMM_PAIR=$m timeout 300 ncu --set full --clock-control none -k regex:conv_igemm_kernel -s 4 -c 2 -o gpurun_out/pairc2_$m -f python scratch/pair_one.py > gpurun_out/ncu_pair$m.log 2>&1; echo "ncu pair$m exit $?"
done
ls -la gpurun_out/*.ncu-rep
