import sys; sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/lf-vio_b200')
import numpy as np
import lfvio
from lfvio.config import PAL_UP, tracker_defaults
from lfvio import synth
from oracle import ransac as R
from oracle.frontend import FeatureTrackerOracle
seed = 0xC0FFEE
cfg = dict(tracker_defaults(), ransac=1, ransac_seed=seed)
src = synth.FrameSource(PAL_UP["image_height"], PAL_UP["image_width"], seed=5)
tr = lfvio.FeatureTracker(PAL_UP, cfg, device=0)
orc = FeatureTrackerOracle(PAL_UP, tracker_defaults(), backend="cv2", ransac=R.make_injected(seed))
for k in range(4):
    img = src.frame(k); pub = k % 3 != 2
    tr.read_image(img, 0.1*k, pub); orc.read_image(img, 0.1*k, pub); tr.update_ids(); orc.update_ids()
    s = orc.snapshot()
    print(k, len(tr.ids), len(s["ids"]), np.array_equal(tr.ids, s["ids"]), np.array_equal(tr.track_cnt, s["track_cnt"]))
    if len(tr.ids) == len(s["ids"]):
        d = np.where(np.any(tr.cur_pts != s["cur_pts"], axis=1))[0]
        print("  differing rows", d[:10], len(d))
        for i in d[:5]: print("   ", i, tr.cur_pts[i], s["cur_pts"][i], tr.ids[i], tr.track_cnt[i])
