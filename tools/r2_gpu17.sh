gcc -O2 -pthread tools/first_touch_probe.c -o /tmp/ftp || exit 1
nproc; cat /sys/kernel/mm/transparent_hugepage/enabled /sys/kernel/mm/transparent_hugepage/defrag; grep -i "MemFree\|MemAvail\|HugePages_Total" /proc/meminfo; numactl -H 2>/dev/null | head -5; lscpu | grep -i "numa\|model name\|socket" | head
for t in 16 32 64; do /tmp/ftp 16 $t 0 1; /tmp/ftp 16 $t 1 1; done
/tmp/ftp 16 16 0 0; /tmp/ftp 16 32 1 0; /tmp/ftp 16 32 2 1
