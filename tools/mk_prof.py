import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import binius_b200
from binius_b200.merkle import BinaryMerkleTree
hal = binius_b200.B200Layer(0)
dev = hal.dev_alloc(1 << 24)
hal.fill(dev, 0x123456789ABCDEF0FEDCBA9876543211)
for _ in range(3):
    t = BinaryMerkleTree.build(hal, dev, 16)
    hal.sync()
