#!/bin/bash
# Round-2 GPU call: GPU test-suite (+ parity report), bench line (ours + reference arm), ncu launch list,
# ncu --set full of one C3 frame and one C4 frame (incl. the L2 miss breakdown by source unit).
# usage: bash scripts/gpu_r2.sh <tag> [tests|notests]      (outputs: gpurun_out/<tag>_*)
T=${1:-r2x}
MODE=${2:-tests}
mkdir -p gpurun_out
rm -f gpurun_out/parity_report.jsonl
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/${T}_smi.txt 2>&1
nproc >> gpurun_out/${T}_smi.txt
if [ "$MODE" = "tests" ]; then
  ( timeout 2400 python -m pytest tests -m gpu -q --durations=15 2>&1 | tail -60 ) > gpurun_out/${T}_pytest.log
  cp gpurun_out/parity_report.jsonl gpurun_out/${T}_parity.jsonl 2>/dev/null
fi
python bench.py --steps 10 --warmup 3 > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${T}_bench_ref.json 2>> gpurun_out/${T}_bench.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/${T}_launches.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline --extra none > gpurun_out/${T}_launches.log 2>&1
XM=lts__t_sectors_srcunit_tex_op_read_lookup_miss.sum,lts__t_sectors_srcunit_tex_op_read_lookup_hit.sum,lts__t_sectors_srcunit_tex_op_write_lookup_miss.sum,lts__t_sectors_srcunit_tex_op_red_lookup_miss.sum,lts__t_sectors_srcunit_tex_lookup_miss.sum,lts__t_sectors_srcunit_ltcfabric_lookup_miss.sum,lts__t_sectors_srcunit_ltcfabric.sum,lts__t_sectors_lookup_miss.sum,lts__t_sectors.sum,l1tex__t_sectors_pipe_lsu_mem_local_op_ld.sum,l1tex__t_sectors_pipe_lsu_mem_local_op_st.sum,l1tex__t_sectors_pipe_lsu_mem_local_op_ld_lookup_miss.sum,l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum,l1tex__t_sectors_pipe_lsu_mem_global_op_ld_lookup_miss.sum,lts__t_sectors_srcunit_tex_aperture_device_lookup_miss.sum,lts__t_sectors_srcunit_tex_aperture_peer_lookup_miss.sum,l1tex__data_pipe_lsu_wavefronts.sum,l1tex__data_pipe_lsu_wavefronts_mem_lg.sum,smsp__inst_executed.sum
timeout 900 ncu --set full --metrics $XM --clock-control none --import-source on -k regex:"trace_kernel|shade_kernel|tail_kernel" -s 7 -c 7 \
    -o gpurun_out/${T}_prof_c3 -f python scripts/exp_c3.py C3 3 > gpurun_out/${T}_ncu_c3.log 2>&1
timeout 900 ncu --set full --metrics $XM --clock-control none --import-source on -k regex:"trace_kernel|shade_kernel|tail_kernel" -s 3 -c 3 \
    -o gpurun_out/${T}_prof_c4 -f python scripts/exp_c3.py C4 2 > gpurun_out/${T}_ncu_c4.log 2>&1
for cfg in C3 C4 C2 C1; do
  echo "=== $cfg" >> gpurun_out/${T}_configs.log
  timeout 300 python scripts/exp_c3.py $cfg 6 >> gpurun_out/${T}_configs.log 2>&1
done
tail -5 gpurun_out/${T}_pytest.log; cut -c1-600 gpurun_out/${T}_bench.json; cut -c1-300 gpurun_out/${T}_bench_ref.json; grep -E "===|frame 5|wave " gpurun_out/${T}_configs.log
