#!/usr/bin/env python
"""Build the tile atlas used by the RGB observation wrappers FROM THE REFERENCE'S OWN
RASTERISER, so that the GPU gather is pixel-exact: for every cell encoding (type, colour, state) and every
(agent_dir, highlight) variant the wrappers can produce, call Grid.render_tile (minigrid.py:475-525) and store
the tile_size x tile_size x 3 result.

    python oracle/gen_atlas.py            # writes gym_minigrid_b200/data/tile_atlas_t8.npz  (needs /root/reference)

atlas uint8 [231][10][T][T][3]; variants: 0 plain, 1 highlighted, 2..5 agent facing dir 0..3 (no highlight),
6 agent facing up (dir 3) + highlight (the agent's own cell of the partial view, minigrid.py:1383-1398),
7..9 agent facing dir 0..2 + highlight (env.render(highlight=True), minigrid.py:1400-1466).
Index 231*... rows for type 10 ('agent') stay zero: that type never reaches a renderer.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import ref_shim as R  # noqa: E402

VARIANTS = [(None, False), (None, True), (0, False), (1, False), (2, False), (3, False), (3, True), (0, True), (1, True), (2, True)]


def main(tile=8):
    R.load_reference()
    mg = sys.modules["gym_minigrid.minigrid"]
    atlas = np.zeros((231, len(VARIANTS), tile, tile, 3), np.uint8)
    valid = np.zeros(231, np.uint8)           # 0: the reference itself cannot render this object (Floor.render uses a removed API, minigrid.py:195-205)
    for t in range(10):                       # unseen .. lava
        for c in range(7):
            for s in range(3):
                if t in (0, 1):
                    obj = None                # unseen / empty decode to None (minigrid.py:124-125)
                else:
                    obj = mg.WorldObj.decode(t, c, s)
                    if obj.type == "goal":
                        obj.color = mg.IDX_TO_COLOR[c]      # real Goal objects keep their colour in the full render
                try:
                    for v, (adir, hl) in enumerate(VARIANTS):
                        mg.Grid.tile_cache.clear()          # the cache is keyed on encode(): never let it alias
                        atlas[t * 21 + c * 3 + s, v] = mg.Grid.render_tile(obj, agent_dir=adir, highlight=hl, tile_size=tile)
                    valid[t * 21 + c * 3 + s] = 1
                except AttributeError:
                    atlas[t * 21 + c * 3 + s] = 0
    out = os.path.join(os.path.dirname(HERE), "gym_minigrid_b200", "data", "tile_atlas_t%d.npz" % tile)
    np.savez_compressed(out, atlas=atlas, tile=np.int32(tile), valid=valid)
    print(out, atlas.shape, "valid codes", int(valid.sum()), "%.1f KB" % (os.path.getsize(out) / 1024))


if __name__ == "__main__":
    main(int(sys.argv[1]) if len(sys.argv) > 1 else 8)
