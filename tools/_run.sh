ncu --set full --clock-control none --import-source on -k regex:inflate_lz_kernel -c 1 -f -o gpurun_out/r45_lz python tools/sweep_inflate.py --streams 65536 --cfgs=-2,14 --steps 1 > gpurun_out/r45_ncu.log 2>&1
tail -3 gpurun_out/r45_ncu.log
