"""Constructor arguments of the modules behind tests/golden/denoise_golden.npz (kept in one place for the generator and the tests)."""
AGG_CFG = dict(in_channel=[16, 24], mid_channel=[8, 16], out_channel=[24, 12], layer_name=['layer1', 'layer2'], rdb_blocks=[1, 2],
               rdb_channel_growth=[8, 8], taf_embs=[2, 3], downsample=[True, False], with_rdb=[True, True], with_taf=[True, True])
