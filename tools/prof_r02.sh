# ncu --set full captures of the split kernels inside a late sweep of the bench workload (round 2).  Every capture is bounded.
# usage (GPU box): bash tools/prof_r02.sh
set -x
timeout 120 python tools/trace_late_sweep.py 30 256 12 notrace > gpurun_out/p_plain.out 2>&1 || exit 1
timeout 500 ncu --set full --clock-control none --import-source on -k regex:jacobi_gram_kernel --launch-skip 150 -c 1 -o gpurun_out/r02_jacobi_gram -f python tools/trace_late_sweep.py 30 256 12 notrace > gpurun_out/p_ncu1.out 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:gs_select_reorth_kernel --launch-skip 1500 -c 1 -o gpurun_out/r02_gs_select_reorth -f python tools/trace_late_sweep.py 30 256 12 notrace > gpurun_out/p_ncu2.out 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:gs_panel_cluster_kernel --launch-skip 1500 -c 1 -o gpurun_out/r02_gs_panel_cluster -f python tools/trace_late_sweep.py 30 256 12 notrace > gpurun_out/p_ncu3.out 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:gs_update_mma_kernel --launch-skip 1500 -c 1 -o gpurun_out/r02_gs_update_mma -f python tools/trace_late_sweep.py 30 256 12 notrace > gpurun_out/p_ncu4.out 2>&1
timeout 120 python tools/probe_mv.py 256 3 20 > gpurun_out/p_mv_plain.out 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:matvec_ --launch-skip 10 -c 2 -o gpurun_out/r02_matvec_tma -f python tools/probe_mv.py 256 3 20 > gpurun_out/p_ncu5.out 2>&1
for f in gpurun_out/p_ncu1.out gpurun_out/p_ncu2.out gpurun_out/p_ncu3.out gpurun_out/p_ncu4.out gpurun_out/p_ncu5.out; do tail -n 3 $f; done; ls -la gpurun_out/*.ncu-rep
