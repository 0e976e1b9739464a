python profiles/scripts/backbone_once.py 16 || exit 1
ncu --set full --clock-control none --import-source on -k regex:k_conv_umma -c 19 -o gpurun_out/r02_backbone -f python profiles/scripts/backbone_once.py 16 > gpurun_out/r02_ncu_backbone.log 2>&1
tail -3 gpurun_out/r02_ncu_backbone.log
