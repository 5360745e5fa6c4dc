#!/bin/bash
# round-2 evidence run (under gpurun, ONE GPU): compute-sanitizer over the small-shape kernel tests, the ncu launch list of
# one eager bench step, `ncu --set full` of the dominant kernels.  Outputs under gpurun_out/ (summaries are copied to
# profiles/ by hand).
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
SAN_TESTS='tests/test_gpu_kernels.py::test_conv_fprop_dgrad_wgrad tests/test_gpu_kernels.py::test_conv_rowfold_stem tests/test_gpu_kernels.py::test_bn_forward_backward tests/test_gpu_kernels.py::test_pool tests/test_gpu_kernels.py::test_pool_inv_matches_reference_kernels tests/test_gpu_kernels.py::test_sparse_sample_fwd_bwd tests/test_gpu_kernels.py::test_build_samples_vs_oracle tests/test_gpu_kernels.py::test_build_samples_centre_corners tests/test_gpu_kernels.py::test_build_samples_clustering tests/test_gpu_kernels.py::test_device_targets_bit_exact tests/test_gpu_kernels.py::test_solver_update tests/test_gpu_kernels.py::test_softmax_nll tests/test_gpu_kernels.py::test_detect_cost tests/test_gpu_kernels.py::test_dgrad_with_fused_bn_backward_statistics tests/test_gpu_inference.py::test_detect_outputs_vs_restatement tests/test_gpu_inference.py::test_detections_nms_vs_reference_golden'
timeout 1200 compute-sanitizer --tool memcheck --error-exitcode 7 --print-limit 20 python -m pytest $SAN_TESTS -q -x -p no:cacheprovider > gpurun_out/r2_sanitizer_memcheck.log 2>&1
echo "memcheck rc=$?" | tee -a gpurun_out/r2_sanitizer_memcheck.log
timeout 500 compute-sanitizer --tool racecheck --error-exitcode 7 --print-limit 20 python -m pytest "tests/test_gpu_kernels.py::test_bn_forward_backward" "tests/test_gpu_kernels.py::test_build_samples_vs_oracle" "tests/test_gpu_kernels.py::test_pool" "tests/test_gpu_kernels.py::test_softmax_nll" -q -x -p no:cacheprovider > gpurun_out/r2_sanitizer_racecheck.log 2>&1
echo "racecheck rc=$?" | tee -a gpurun_out/r2_sanitizer_racecheck.log
timeout 400 compute-sanitizer --tool racecheck --error-exitcode 7 --print-limit 40 python -m pytest "tests/test_gpu_kernels.py::test_conv_fprop_dgrad_wgrad" -q -k "False and case0" -p no:cacheprovider > gpurun_out/r2_sanitizer_racecheck_conv.log 2>&1
echo "racecheck conv rc=$?" | tee -a gpurun_out/r2_sanitizer_racecheck_conv.log
# launch list of one eager step
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 2500 --csv --log-file gpurun_out/r2_launches.csv python bench.py --steps 1 --warmup 1 --eager --no-cpu-baseline --no-parity-mode > gpurun_out/r2_ncu_bench.log 2>&1
echo "launch list rc=$?"
# full sections of the dominant kernels, one steady-state step (skip the first 600 matching launches = 2+ warm-up steps)
K='regex:conv_fprop_halo2|conv_fprop_halo_kernel|conv_fprop_kernel|conv_wgrad|bn_bwd_fused|bn_apply_kernel|pool_inv|sparse_sample|corner_select|pair_select|wgrad_reduce_multi|maxpool|solver_update|weight_prep_multi'
timeout 2400 ncu --set full --clock-control none -k "$K" -s 700 -c 240 -o gpurun_out/r2_full python bench.py --steps 1 --warmup 1 --eager --no-cpu-baseline --no-parity-mode > gpurun_out/r2_ncu_full.log 2>&1
echo "ncu full rc=$?"
python scripts/ncu_summary.py gpurun_out/r2_full.ncu-rep > gpurun_out/r2_ncu_full_summary.txt 2>&1
ls -la gpurun_out/r2_full.ncu-rep
# the report itself is scratch (tens of MB): keep the summary, drop the report when it would not fit the 64 MiB return path
if [ $(stat -c %s gpurun_out/r2_full.ncu-rep) -gt 30000000 ]; then rm -f gpurun_out/r2_full.ncu-rep; fi
tail -5 gpurun_out/r2_ncu_full_summary.txt
du -sh gpurun_out
