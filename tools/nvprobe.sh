cd /root/repo
for b in 74 148 256; do echo "== B=$b"; python tools/netvlad_timeline.py $b 64 2>&1 | grep -E "video 0" | grep -E "p0 loads issued|a_ready seen|last group committed|rescale done"; done
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,lts__t_sector_hit_rate.pct,lts__t_bytes.sum -k regex:netvlad --csv python tools/netvlad_timeline.py 256 64 2>/dev/null | grep -E "netvlad" | awk -F'","' '{print $(NF-2), $(NF-1), $NF}' | tail -20
ncu --metrics dram__bytes_read.sum,gpu__time_duration.sum,lts__t_sector_hit_rate.pct -k regex:netvlad --csv python tools/netvlad_timeline.py 74 64 2>/dev/null | grep -E "netvlad" | awk -F'","' '{print $(NF-2), $(NF-1), $NF}' | tail -12
