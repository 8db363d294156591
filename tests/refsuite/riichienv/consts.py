# tile-dimension constants of src/riichienv/consts.py (values are facts of the game)
N_TILE_TYPES_4P = 34
N_TILE_TYPES_3P = 27
N_TILES_4P = 136
N_TILES_3P = 108
