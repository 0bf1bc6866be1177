"""Camera-visibility culling on the GPU: the device
result equals the CPU restatement that tests/golden/mark_untrained_reference.npz pins, and a density-grid update after
the culling only ever touches trainable cells."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def test_mark_untrained_density_grid_on_device_and_update_respects_it(oracle):
    from jaxngp_b200 import ogrid, synthetic as S
    from oracle import ogrid_np
    dev = "cuda:0"
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "mark_untrained_reference.npz"))
    G, K, bound = int(g["G"]), int(g["K"]), float(g["bound"])
    G3 = G ** 3
    cam = S.camera()
    for step in (0, 300):
        grid = ogrid.OccupancyDensityGrid(K, G, device=dev)
        grid.density.copy_(torch.from_numpy(g["density_in"]).to(dev))
        alive = ogrid.mark_untrained_density_grid(grid, torch.from_numpy(g["poses"]).to(dev), cam, bound, 1024, step)
        assert np.array_equal(grid.density.cpu().numpy(), g[f"step{step}_density"])
        assert np.array_equal(grid.occ_mask.cpu().numpy(), g[f"step{step}_occ_mask"])
        assert np.array_equal(grid.occupancy.cpu().numpy(), g[f"step{step}_occupancy"])
        assert np.array_equal(grid.alive_indices.cpu().numpy().astype(np.uint32), g[f"step{step}_alive_indices"])
        assert grid.alive_indices_offset == g[f"step{step}_alive_indices_offset"].tolist()
        assert np.array_equal(alive.cpu().numpy(), ogrid_np.visible_cells(K, G, bound, g["poses"], cam))
    # updates after the culling: culled cells keep -1 whatever is sampled
    culled = grid.density < 0
    gen = torch.Generator(device=dev).manual_seed(0)
    for update_all in (True, False):
        for cas in range(K):
            idx, _, _ = ogrid.update_ogrid_density(grid, lambda xyz: torch.full((xyz.shape[0],), 3.0, device=dev), cas, update_all,
                                                   bound, 1 << 16, generator=gen)
            alive_c = set(grid.alive_in_cascade(cas).cpu().tolist())
            assert set(idx.cpu().tolist()) <= alive_c
            if update_all:
                assert idx.shape[0] == len(alive_c)
    assert torch.equal(grid.density < 0, culled)
