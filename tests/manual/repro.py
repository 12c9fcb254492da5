import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import sourmash_rust_b200 as smb
a = smb.KmerMinHash(20, 10)
a.add_sequence(b"TGCCGCCCAGCA")
print(a.mins)
