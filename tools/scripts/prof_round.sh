python tools/prof_conv.py f16 192 192 && ncu --set full --clock-control none --import-source on -k regex:k_conv_tc -s 2 -c 1 -o gpurun_out/conv_f16_192 -f python tools/prof_conv.py f16 192 192 > gpurun_out/prof_c192.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_conv_tc -s 2 -c 1 -o gpurun_out/conv_f16_64 -f python tools/prof_conv.py f16 64 64 > gpurun_out/prof_c64.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_pointnet_tc -s 1 -c 1 -o gpurun_out/pointnet_tc_r1 -f python tools/prof_pointnet.py > gpurun_out/prof_pntc.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -s 1500 -c 1300 --csv --log-file gpurun_out/launches_r1_f16.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --eager > gpurun_out/ncu_eager.log 2>&1
wc -l gpurun_out/launches_r1_f16.csv
