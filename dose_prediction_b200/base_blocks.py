"""Drop-in for the hot-path classes of `OARSegmentation/Models/Nets/base_blocks.py` (MultiUnetBasicBlock :12,
ModifiedUnetrUpBlock :91, ModifiedUnetOutBlock :144) and `blocks_MDUNet.py` (conv_block_3 :64, conv_block_7 :98,
conv_3_1 :132, DualDilatedBlock :194)."""
from .networks import (DualDilatedBlock, ModifiedUnetOutBlock, ModifiedUnetrUpBlock, MultiUnetBasicBlock,  # noqa: F401
                       conv_3_1, conv_block_3, conv_block_7)

__all__ = ["MultiUnetBasicBlock", "ModifiedUnetrUpBlock", "ModifiedUnetOutBlock", "conv_block_3", "conv_block_7", "conv_3_1",
           "DualDilatedBlock"]
