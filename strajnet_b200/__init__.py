"""strajnet_b200: B200-native (sm_100a CUDA behind a C ABI) occupancy-flow forward path with the
call surface of georgeliu233/STrajNet's Keras layers.  See DESIGN.md / INTEGRATION.md."""
from .layers import (BasicLayer, Dense, FGMSA, Mlp, PatchEmbed, PatchMerging, Pyramid3DDecoder, STrajNet,
                     SwinTransformerBlock, SwinTransformerEncoder, TrajNetCrossAttention, WindowAttention,
                     relative_position_index, shift_attn_mask, window_partition, window_reverse, window_token_map)

__all__ = ["BasicLayer", "Dense", "FGMSA", "Mlp", "PatchEmbed", "PatchMerging", "Pyramid3DDecoder", "STrajNet",
           "SwinTransformerBlock", "SwinTransformerEncoder", "TrajNetCrossAttention", "WindowAttention",
           "relative_position_index", "shift_attn_mask", "window_partition", "window_reverse", "window_token_map"]
