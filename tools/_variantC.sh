set -e
cd /root/repo/node-cueball_b200/csrc
sed -i 's/constexpr int TILE = 2048; /constexpr int TILE = 1024; /; s/constexpr int TPB = 256; /constexpr int TPB = 128; /' cb_device.cuh
sed -i 's/__launch_bounds__(TPB, 4) scan_kernel/__launch_bounds__(TPB, 8) scan_kernel/' cb_scan.cuh
grep -n "constexpr int TILE\|constexpr int TPB" cb_device.cuh
make > /dev/null 2>&1 || (tail -5 build.log; exit 1)
grep -A2 "scan_kernelILb1ELb0" build.log | grep -E "spill|Used"
