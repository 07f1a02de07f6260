#!/bin/bash
# steady-state DRAM traffic of the group K1 launch under different L2 pin settings (no cache flush between launches)
for pin in 2 4 6; do
  echo "== pin_cams=$pin"
  ESVIO_K1_PIN_CAMS=$pin ncu --cache-control none --clock-control none --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,lts__t_sector_hit_rate.pct -k regex:k_sae_update_ts -s 10 -c 1 --csv python scratch/group_k1.py stereo_vga_5mevs 8 2>&1 | grep -E "dram__bytes|gpu__time|hit_rate" | awk -F'","' '{print $(NF-2), $(NF-1), $NF}'
done
