set -x
python bench.py --steps 2 --warmup 3 --no-cpu-baseline --eager > gpurun_out/plain_eager.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -s 6000 -c 3200 --csv --log-file gpurun_out/launches_r1_f16.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --eager > gpurun_out/ncu_eager.log 2>&1
wc -l gpurun_out/launches_r1_f16.csv
python tools/prof_conv.py f16 192 192 && ncu --set full --clock-control none --import-source on -k regex:k_conv_tc -s 2 -c 1 -o gpurun_out/conv_f16_192 -f python tools/prof_conv.py f16 192 192 > gpurun_out/prof_c192.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_conv_tc -s 2 -c 1 -o gpurun_out/conv_f16_64 -f python tools/prof_conv.py f16 64 64 > gpurun_out/prof_c64.log 2>&1
python tools/prof_pointnet.py && ncu --set full --clock-control none --import-source on -k regex:k_pointnet_tc -s 1 -c 1 -o gpurun_out/pointnet_tc_r1 -f python tools/prof_pointnet.py > gpurun_out/prof_pntc.log 2>&1
python tools/prof_hbm.py && ncu --set full --clock-control none -k regex:"k_im2row|k_slice|k_splat|k_distribute_rows|k_insert_points" -s 5 -c 5 -o gpurun_out/hbm_r1b -f python tools/prof_hbm.py > gpurun_out/prof_hbm.log 2>&1
ls -la gpurun_out/*.ncu-rep
