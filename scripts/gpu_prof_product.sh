#!/bin/bash
# one full ncu capture of the product kernel alone (C3)
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"qp_product" -s 1 -c 1 -f -o gpurun_out/prof_product_c3 python bench.py --problems 20000 --steps 1 --warmup 1 --no-e2e --no-cpu --no-configs --no-workspace --no-weak > /dev/null 2>&1
python scripts/ncu_summary.py gpurun_out/prof_product_c3.ncu-rep 30 > gpurun_out/ncu_product_c3_summary.txt 2>&1
cat gpurun_out/ncu_product_c3_summary.txt
ncu -i gpurun_out/prof_product_c3.ncu-rep --page raw --csv 2>/dev/null | python -c "
import csv,sys
rows=list(csv.reader(sys.stdin))
h=rows[0]; v=rows[2] if len(rows)>2 else rows[1]
want=['dram__bytes_read.sum','dram__bytes_write.sum','dram__throughput.avg.pct_of_peak_sustained_elapsed','l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum','sm__warps_active.avg.pct_of_peak_sustained_active','launch__occupancy_limit_registers','launch__occupancy_limit_shared_mem','smsp__cycles_active.avg','l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum','l1tex__t_requests_pipe_lsu_mem_global_op_st.sum','l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum','l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum','lts__t_sectors_op_write.sum','lts__t_sectors_op_read.sum']
for k in want:
    for i,n in enumerate(h):
        if n==k: print(k, v[i])
"
