"""Integer-pipe microbenchmarks on the device (roofline denominators of the sketch kernel).
Modes 0-2: dependent chains (legacy); 3-9: eight independent chains per pipe, SASS counted with cuobjdump."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import sourmash_rust_b200 as smb
dev, sms = smb.device_info()
# SASS instructions per inner step (per 8-chain sweep); smgpu_int_peak reports 8 per sweep for modes 0-2
# and we pass iters so that its "64 per iteration" bookkeeping is rescaled here
spec = {3: ("8 LOP3", 8), 4: ("8 SHF", 8), 5: ("8 IMAD + 8 LOP3", 16), 6: ("8 IMAD.WIDE + 8 LOP3", 16),
        7: ("4 IMAD.WIDE + 8 IMAD + 12 LOP3/SHF", 24), 8: ("8 IMAD.HI", 8), 9: ("8 IMAD + 8 SHF", 16)}
for mode, (what, per_sweep) in spec.items():
    for blocks in (sms * 4, sms * 8):
        # the C side counts iters * 64 * 256 * blocks "instructions"; real = iters * 4 sweeps * per_sweep
        r = smb.int_peak(mode, 4096, blocks) * (4 * per_sweep) / 64.0
        print("mode %d (%s) blocks %d: %.2f T thread-instr/s = %.1f thread-instr/clk/SM at 1965 MHz" %
              (mode, what, blocks, r / 1e12, r / 1.965e9 / sms))

# hash-only ceiling: smgpu_int_peak counts iters * 64 * 256 * blocks "instructions" = hashes
for mode, k in ((10, 21), (11, 31), (12, 51)):
    r = smb.int_peak(mode, 256, sms * 8)
    print("mode %d (MurmurHash3 x64_128 of %d register-resident bytes only): %.1f G hashes/s" % (mode, k, r / 1e9))
