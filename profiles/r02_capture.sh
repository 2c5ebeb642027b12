#!/bin/bash
# Round-2 profile capture (run on the GPU box through gpurun): launch list of bench.py, `ncu --set full` of the
# aggregation kernel (base graph and 8x graph) and of the step's dense / BatchNorm kernels.  The .ncu-rep files are
# exported to CSV pages on the box and only the two single-launch aggregation reports are kept (gpurun_out <= 64 MiB).
set -u
O=gpurun_out/r02
mkdir -p $O
python bench.py --steps 3 --warmup 3 > $O/plain_bench.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file $O/launches_r02.csv \
    python bench.py --steps 3 --warmup 3 > $O/ncu_bench2.log 2>&1
exp() {  # exp <name>: raw + details pages of $O/<name>.ncu-rep as text
  ncu -i $O/$1.ncu-rep --page raw --csv > $O/$1.raw.csv 2>/dev/null
  ncu -i $O/$1.ncu-rep --page details > $O/$1.details.txt 2>/dev/null
}
for shape in 168 64; do
  python profiles/spmm_ncu_probe.py $shape > $O/plain_probe.log 2>&1 &&
  ncu --set full --clock-control none --import-source on -k regex:spmm_stream -s 2 -c 1 -f -o $O/r02_spmm_x1_F$shape \
      python profiles/spmm_ncu_probe.py $shape > $O/ncu_probe1.log 2>&1
  exp r02_spmm_x1_F$shape
done
for shape in 168 128 64; do
  python profiles/spmm_ncu_probe_x8.py $shape > $O/plain_probe8.log 2>&1 &&
  ncu --set full --clock-control none --import-source on -k regex:spmm_stream -s 1 -c 1 -f -o $O/r02_spmm_x8_F$shape \
      python profiles/spmm_ncu_probe_x8.py $shape > $O/ncu_probe8.log 2>&1
  exp r02_spmm_x8_F$shape
done
python profiles/step_ncu_probe.py > $O/plain_step.log 2>&1 &&
ncu --set full --clock-control none -k regex:"gemm_tn_kernel|gemm_wgrad_kernel|bn_act_fwd8|bn_relu_bwd" -s 33 -c 11 -f \
    -o $O/r02_step_kernels python profiles/step_ncu_probe.py > $O/ncu_step.log 2>&1
exp r02_step_kernels
rm -f $O/r02_step_kernels.ncu-rep $O/r02_spmm_x1_F64.ncu-rep $O/r02_spmm_x8_F128.ncu-rep $O/r02_spmm_x8_F64.ncu-rep
ls -la $O; du -sh $O
