"""GPU: N4 -- on-device disparity metrics and the KITTI 16-bit encoding against the oracle's restatement of
nmrf/utils/evaluation.py:326-359 and nmrf/utils/frame_utils.py:237-239."""
import numpy as np
import pytest
import torch

from oracle import nmrf_oracle as O

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("only_valid,max_disp", [(True, 192), (False, 192), (True, None)])
def test_disp_evaluator_matches_reference_metrics(only_valid, max_disp):
    from nmrf_b200.evaluation import DispEvaluator
    g = torch.Generator().manual_seed(5)
    B, H, W = 5, 93, 211                                             # ragged: HW not a multiple of the block's 4096 pixels
    gt = torch.rand(B, H, W, generator=g) * 250 + 0.5
    pr = gt + torch.randn(B, H, W, generator=g) * 2.5
    pr[0, :10] += 40                                                 # some gross outliers
    valid = torch.rand(B, H, W, generator=g) > 0.3
    valid[3] = False                                                 # an image without a valid pixel (skipped when only_valid)
    gt[4] = 1000.0                                                   # ... and one entirely beyond max_disp
    thres = ["0.5", "1.0", "3.0"]
    ev = DispEvaluator(thres, only_valid, max_disp)
    # two `process` calls (3 + 2 images), as a data loader would deliver them
    for sl in (slice(0, 3), slice(3, 5)):
        ev.process({"disp": gt[sl], "valid": valid[sl]}, {"disp": pr[sl].cuda()})
    got = ev.evaluate()["disp"]
    ref = O.disp_metrics(pr, gt, valid, only_valid, np.inf if max_disp is None else max_disp, thres)
    assert set(got) == set(ref)
    for k in ref:
        assert abs(got[k] - ref[k]) <= 2e-5 * max(1.0, abs(ref[k])), (k, got[k], ref[k])


def test_evaluator_rejects_eval_prop():
    from nmrf_b200.evaluation import DispEvaluator
    with pytest.raises(NotImplementedError, match="superpixel"):
        DispEvaluator(None, True, eval_prop=True)


def test_kitti_u16_encoding_bit_exact(tmp_path):
    from nmrf_b200.evaluation import disp_to_kitti_u16, write_disp_kitti
    g = torch.Generator().manual_seed(6)
    disp = torch.rand(375, 1242, generator=g) * 255.9
    disp[0, :8] = torch.tensor([0.0, 0.001953125, 0.005859375, 1.5 / 256, 2.5 / 256, 255.998046875, 100.0, 17.3])   # exact .5 ties
    got = disp_to_kitti_u16(disp.cuda()).cpu().numpy()
    assert got.dtype == np.uint16 and np.array_equal(got, O.kitti_u16(disp))
    path = str(tmp_path / "000000_10.png")
    write_disp_kitti(path, disp.cuda())
    import cv2
    back = cv2.imread(path, cv2.IMREAD_ANYDEPTH)
    assert back.dtype == np.uint16 and np.array_equal(back, O.kitti_u16(disp))
