# Round-2 profile refresh (run on the GPU box through gpurun): launch lists of one cfg5 chunk and one cfg2 sweep, and `--set full`
# captures of the DMMA GEMM tiles, condensed on the box (the .ncu-rep files are too large to bring back together).
set -x
python tools/cfg5_chunk.py 1 > gpurun_out/c5.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_cfg5b.csv python tools/cfg5_chunk.py 1 > gpurun_out/c5n.log 2>&1
python tools/cfg5_chunk.py 1 > gpurun_out/c5.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:gemm_kernel -s 40 -c 4 -o gpurun_out/prof_gemm_c128b_r02 -f python tools/cfg5_chunk.py 1 > gpurun_out/c5f.log 2>&1
python tools/ncu_summary.py gpurun_out/prof_gemm_c128b_r02.ncu-rep > gpurun_out/ncu_gemm_c128_r02.txt 2>&1; rm -f gpurun_out/prof_gemm_c128b_r02.ncu-rep
python tools/run_matvec.py 2 > gpurun_out/mv.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:gemm_ -c 4 -o gpurun_out/prof_gemm_f64b_r02 -f python tools/run_matvec.py 2 > gpurun_out/mvf.log 2>&1
python tools/ncu_summary.py gpurun_out/prof_gemm_f64b_r02.ncu-rep > gpurun_out/ncu_gemm_f64_r02.txt 2>&1; rm -f gpurun_out/prof_gemm_f64b_r02.ncu-rep
python tools/cfg2_sweep.py 1 > gpurun_out/c2.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_cfg2b.csv python tools/cfg2_sweep.py 1 > gpurun_out/c2n.log 2>&1
ls -la gpurun_out/
