import importlib, numpy as np, sys
import os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
v=importlib.import_module('vision-enhanced-lidar-odometry_b200')
api,syn=v.api,v.synth
P,Tr,w,h=syn.calib_raw(0); cal=api.calib_from_kitti(P,Tr,w,h)
prm=api.default_params(max_slots=2)
c=api.Context(prm,cal)
a,_=syn.scan(1000); b,_=syn.scan(1001)
c.scan_upload(0,a); c.scan_upload(1,b)
for it,p in ((1,0),(1,2),(2,3),(2,5)):
    corr,neq,kept=c.icp_pass(1,0,syn.pose_guess(1001,p),it,1)
    q=neq[58]
    print('iter',it,'pass',p,'q',q,'kept',kept,'seed/q %.1f exh/q %.1f rings/q %.2f mask/q %.2f'%(neq[59]/q,neq[60]/q,neq[61]/q,neq[62]/q))
    k=corr['kept']==1
    d=np.linalg.norm(corr['v0'][k],axis=1)
    print('   residual abs median %.4f'%np.median(np.abs(corr['residual'][k])))
