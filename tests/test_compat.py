"""Caller-side shims (esac_b200/compat.py): the dataset tuple and the shift augmentation of the reference's loops."""
import random

import torch

from esac_b200.compat import OUTPUT_SUBSAMPLE, SyntheticRoomDataset, random_shift


def test_dataset_yields_the_reference_six_tuple():
    ds = SyntheticRoomDataset(num_experts=3, length=4, hypotheses=32, seed=1)
    loader = torch.utils.data.DataLoader(ds, shuffle=True, num_workers=0)       # train_esac.py:77
    seen = 0
    for idx, image, focallength, gt_pose, gt_coords, gt_expert in loader:
        assert image.shape == (1, 1, 480, 640) and image.dtype == torch.float32
        assert float(focallength[0]) == 525.0
        assert gt_pose.shape == (1, 4, 4) and gt_pose.dtype == torch.float32
        assert gt_coords.shape == (1, 3, 60, 80)
        assert 0 <= int(gt_expert[0]) < 3
        pred = ds.prediction_for(int(idx))
        assert pred.shape == (3, 3, 60, 80)
        assert torch.equal(pred[int(gt_expert[0])], gt_coords[0])
        seen += 1
    assert seen == 4
    a = ds[2]
    b = SyntheticRoomDataset(num_experts=3, length=4, hypotheses=32, seed=1)[2]
    assert torch.equal(a[1], b[1]) and torch.equal(a[3], b[3])                  # deterministic in (seed, index)
    assert SyntheticRoomDataset(training=False)[0][4] == 0                      # room_dataset.py:209-212


def test_random_shift_accepts_the_callers_float_bound_and_shifts_by_padding():
    random.seed(4)
    img = torch.arange(1 * 1 * 6 * 8, dtype=torch.float32).reshape(1, 1, 6, 8) + 1
    for _ in range(20):
        padX, padY, out = random_shift(img, OUTPUT_SUBSAMPLE / 2)               # train_esac.py:125 passes 4.0
        assert isinstance(padX, int) and -4 <= padX <= 4 and -4 <= padY <= 4
        assert out.shape == img.shape
        # content moves right/down by (padX, padY); what slides in is zero
        for y in range(6):
            for x in range(8):
                sy, sx = y - padY, x - padX
                want = img[0, 0, sy, sx] if 0 <= sy < 6 and 0 <= sx < 8 else 0.0
                assert out[0, 0, y, x] == want
