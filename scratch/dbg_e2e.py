import sys, numpy as np
sys.path.insert(0, '.')
from esvio_b200 import synth, frontend
from oracle import oracle as ora
W,H,rate=346,260,1.0e6
use_ransac=int(sys.argv[1]) if len(sys.argv)>1 else 0
cfg=synth.default_config(W,H,use_ransac=use_ransac,max_events_per_window=1<<20)
ft=frontend.FeatureTracker(cfg); ot=ora.OracleTracker(cfg,use_cv2=False,disable_ransac=not use_ransac)
s=synth.StereoEventStream(W,H,rate)
for k in range(12):
    L,R,t=s.stereo_window(k); pub=(k%2)==0
    ft.PUB_THIS_FRAME=pub; ft.trackEvent(t,L,R); o=ot.track(t,L,R,pub)
    common,ia,ib=np.intersect1d(ft.ids,o['id'],return_indices=True)
    d=np.hypot(ft.cur_pts[ia,0]-o['u'][ib],ft.cur_pts[ia,1]-o['v'][ib]) if len(common) else np.zeros(0)
    print(k,pub,'n',len(ft.ids),len(o['id']),'common',len(common),'maxd',d.max() if len(d) else 0,'stats',ft.stats,o['stats'])
    bad=np.nonzero(d>1e-3)[0]
    for b in bad[:6]:
        print('   id',common[b],'gpu',ft.cur_pts[ia[b]],'cpu',o['u'][ib[b]],o['v'][ib[b]],'cnt',ft.track_cnt[ia[b]],o['track_cnt'][ib[b]])
