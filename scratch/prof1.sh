set -e
python tools/prof_pointnet.py
ncu --set full --clock-control none --import-source on -k regex:k_pointnet_mlp_max -s 2 -c 1 -o gpurun_out/pointnet_r1 -f python tools/prof_pointnet.py > gpurun_out/prof_pn.log 2>&1 || tail -5 gpurun_out/prof_pn.log
python tools/prof_conv.py
ncu --set full --clock-control none --import-source on -k regex:k_conv_tc -s 2 -c 1 -o gpurun_out/conv_f16_r1 -f python tools/prof_conv.py f16 > gpurun_out/prof_c16.log 2>&1 || tail -5 gpurun_out/prof_c16.log
ls -la gpurun_out/*.ncu-rep
