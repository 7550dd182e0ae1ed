"""A/B timing of trgl_linear_ls across several builds of the library (experiment helper)."""
import ctypes, sys, os, numpy as np
ROOT=os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0,ROOT+'/multiple-quadrotor-slam_b200')
import synthetic_rig as rig
n=int(sys.argv[1]); libs=sys.argv[2:]
u1b,P1,u2b,P2,_=rig.make_correspondences(2_000_000,'rotating',0.8)
reps=-(-n//len(u1b)); u1=np.tile(u1b,(reps,1))[:n]; u2=np.tile(u2b,(reps,1))[:n]
P1=np.ascontiguousarray(P1); P2=np.ascontiguousarray(P2)
vp=ctypes.c_void_p; dp=ctypes.POINTER(ctypes.c_double)
for path in libs:
    L=ctypes.CDLL(path)
    L.trgl_device_alloc.argtypes=[ctypes.POINTER(vp),ctypes.c_size_t]; L.trgl_memcpy_h2d.argtypes=[vp,vp,ctypes.c_size_t,vp]
    L.trgl_linear_ls.argtypes=[vp,vp,dp,dp,vp,vp,ctypes.c_int64,ctypes.c_int,ctypes.c_int,vp]
    L.trgl_event_create.argtypes=[ctypes.POINTER(vp)]; L.trgl_event_record.argtypes=[vp,vp]; L.trgl_event_elapsed_ms.argtypes=[vp,vp,ctypes.POINTER(ctypes.c_float)]
    L.trgl_device_free.argtypes=[vp]
    def alloc(b):
        p=vp(); assert L.trgl_device_alloc(ctypes.byref(p),b)==0; return p
    d1=alloc(u1.nbytes); d2=alloc(u2.nbytes); x=alloc(n*24); st=alloc(n)
    L.trgl_memcpy_h2d(d1,u1.ctypes.data,u1.nbytes,None); L.trgl_memcpy_h2d(d2,u2.ctypes.data,u2.nbytes,None); L.trgl_device_synchronize()
    cfgs=[('ppt',1),('ppt',2),('ppt',4)]
    cfgs=[(k,v,None) for (k,v) in cfgs]
    if hasattr(L,'trgl_set_stream_variant'): cfgs+=[('var',v,None) for v in (1,2,3,4,5,6)]
    for kind,v,f in cfgs:
        if hasattr(L,'trgl_set_stream_variant'): L.trgl_set_stream_variant(0 if kind=='ppt' else v)
        if kind=='ppt': L.trgl_set_points_per_thread(v)
        ev=[]
        for i in range(14):
            e=vp(); L.trgl_event_create(ctypes.byref(e)); ev.append(e)
        for _ in range(3): rc=L.trgl_linear_ls(d1,d2,P1.ctypes.data_as(dp),P2.ctypes.data_as(dp),x,st,n,0,1,None)
        if L.trgl_device_synchronize()!=0: print(os.path.basename(path),kind,v,'ERROR'); break
        for i in range(13):
            L.trgl_event_record(ev[i],None)
            if i<12: L.trgl_linear_ls(d1,d2,P1.ctypes.data_as(dp),P2.ctypes.data_as(dp),x,st,n,0,1,None)
        L.trgl_device_synchronize()
        ms=[]
        for i in range(12):
            f=ctypes.c_float(); L.trgl_event_elapsed_ms(ev[i],ev[i+1],ctypes.byref(f)); ms.append(f.value)
        m=float(np.median(ms)); print('%-22s %s=%d n=%d  %.4f ms  %.0f GB/s  frac %.3f'%(os.path.basename(path),kind,v,n,m,57*n/m/1e6,57*n/m/1e6/6543.1))
    for p in (d1,d2,x,st): L.trgl_device_free(p)
