cd /root/repo
for cfg in "0 1" "0 0" "1 1" "2 1" "3 1" "7 1"; do set -- $cfg; python tools/netvlad_timeline.py 256 64 $1 $2 2>&1 | grep -E "kernel|video 0" | grep -E "kernel|a_ready seen|group1 ready|group3 ready|last group committed|epilogue done|rescale done"; done
