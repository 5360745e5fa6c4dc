K='regex:conv_wgrad2_kernel|conv_fprop_halo2_kernel|conv_wgrad_kernel|conv_wgrad_rows|bn_bwd_fused|wgrad_reduce_multi'
timeout 900 ncu --set full --clock-control none -k "$K" -s 420 -c 110 -o gpurun_out/r2_full_final python bench.py --steps 1 --warmup 1 --eager --no-cpu-baseline --no-parity-mode > gpurun_out/r2_ncu_full_final.log 2>&1
echo "ncu rc=$?"
python scripts/ncu_summary.py gpurun_out/r2_full_final.ncu-rep > gpurun_out/r2_ncu_full_final.txt 2>&1
ls -la gpurun_out/r2_full_final.ncu-rep
rm -f gpurun_out/r2_full_final.ncu-rep
head -70 gpurun_out/r2_ncu_full_final.txt | cut -c1-170
