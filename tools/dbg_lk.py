import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
from conftest import processed_pair
from test_gpu_lk import _points
from oracle import oracle as O
import lk_optical_flow_method_b200.cv as cv
for (W,H,n,q,d,bs,win,ml,crit) in [(860,483,500,0.01,7,7,(15,15),2,(3,10,0.03)),(1920,1080,2000,0.01,7,7,(21,21),2,(3,10,0.03)),(641,359,300,0.05,3.5,3,(31,31),4,(3,30,0.01))]:
    prev,nxt=processed_pair(W,H)
    pts=_points(O,prev,W,H,n,q,d,bs).reshape(-1,1,2)
    got=cv.calcOpticalFlowPyrLK(prev,nxt,pts,None,winSize=win,maxLevel=ml,criteria=crit)
    ref=O.pyrlk(prev,nxt,pts,win,ml,crit,return_iters=True)
    both=((got[1]==1)&(ref[1]==1)).ravel()
    df=np.abs(got[0]-ref[0]).reshape(-1,2).max(1); de=np.abs(got[2]-ref[2]).ravel()
    print(win,'agree',(got[1]==ref[1]).mean(),'maxdflow',df[both].max(),'maxderr',de[both].max(), 'n>1e-3 flow',(df[both]>1e-3).sum(), 'n err>1e-3',(de[both]>1e-3).sum())
    i=np.argmax(np.where(both,de,0)); print('  worst err pt',i,pts[i].ravel(),'flow diff',df[i],'err',got[2][i],ref[2][i],'iters',ref[3][i])
    try:
        import cv2
        c=cv2.calcOpticalFlowPyrLK(prev,nxt,pts,None,winSize=win,maxLevel=ml,criteria=crit)
        b2=((got[1]==1)&(c[1]==1)).ravel()
        print('  vs cv2: agree',(got[1]==c[1]).mean(),'dflow',np.abs(got[0]-c[0]).reshape(-1,2).max(1)[b2].max(),'derr',np.abs(got[2]-c[2]).ravel()[b2].max(), '| oracle vs cv2 derr', np.abs(ref[2]-c[2]).ravel()[((ref[1]==1)&(c[1]==1)).ravel()].max())
    except ImportError: pass
