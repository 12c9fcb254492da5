import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import sourmash_rust_b200 as smb
dev, sms = smb.device_info()
# counted instructions per unrolled step: mode 0 = 8 IMAD, mode 1 = 4 SHF + 8 LOP3 (two per chain element) , mode 2 = 4 IMAD + 2 SHF + 4 LOP3
real = {0: 8, 1: 12, 2: 10}
for mode in (0, 1, 2):
    for blocks in (sms * 4, sms * 8):
        r = smb.int_peak(mode, 8192, blocks) * real[mode] / 8.0
        print("mode %d blocks %d: %.2f T thread-instr/s = %.1f instr/clk/SM at 1965 MHz" % (mode, blocks, r / 1e12, r / 1.965e9 / sms))
