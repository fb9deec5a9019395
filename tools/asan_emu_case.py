"""Driver of tools/asan_emu.sh (ASan/UBSan over the host-stepped device code)."""
import os, sys, ctypes as C
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'tests')); sys.path.insert(0, ROOT)
import orc, emu, raygen, synth, numpy as np, importlib
# swap in the ASan build
L = C.CDLL(os.environ.get('JT_EMU_LIB', '/tmp/libjt_emu_asan.so'))
for name in ['emu_create','emu_destroy','emu_last_error','emu_stats','emu_intersect','emu_intersect_instance','emu_trace_range','emu_trace_wavefront','emu_wide_counts']:
    pass
emu._LIB = None
real = emu.lib()
for name in ['emu_create','emu_destroy','emu_last_error','emu_stats','emu_intersect','emu_intersect_instance','emu_trace_range','emu_trace_wavefront','emu_wide_counts','emu_set_suspend_every','emu_resumed_rays','emu_stolen_samples']:
    f = getattr(L, name); g = getattr(real, name); f.restype = g.restype; f.argtypes = g.argtypes
emu._LIB = L
bvhm=importlib.import_module('julia-raytracer_b200.bvh'); lm=importlib.import_module('julia-raytracer_b200.lights')
for s in ['synthetic_all','synthetic_one','cornellbox','features1','classroom','ecosys']:
    sc = synth.make_scene(s) if s.startswith('synth') else orc.jt.load_scene(os.path.join(ROOT, 'assets', 'scenes', f'{s}.jtscene'))
    b=bvhm.make_scene_bvh(sc); Lt=lm.make_trace_lights(sc)
    o=orc.Oracle(sc,b,Lt); e=emu.Emu(sc,b,Lt)
    p=orc.make_params(resolution=48, samples=2, batch=2, sampler=1); w,h=o.make_state(p)
    rays=raygen.camera_rays(o,p,w,h,3000,seed=3); sec=raygen.secondary_rays(rays,o.intersect(rays),seed=4)
    for trav in (0,1,3): e.intersect(np.concatenate([rays,sec]),trav)
    for sampler in (1,2):
        for trav in (0,1):
            p=orc.make_params(resolution=32, samples=2, batch=2, sampler=sampler, traversal=trav)
            e.trace(p,32,max(1,int(32/ float(sc.cameras[0].aspect))),0,2,wavefront=True); e.trace(p,32,max(1,int(32/float(sc.cameras[0].aspect))),0,2,wavefront=False)
    # round 2: park / resume of every ray every 2 traversal steps + a longer range with stolen samples
    if s != 'ecosys':
        p=orc.make_params(resolution=32, samples=8, batch=8, sampler=1, traversal=0)
        hh=max(1,int(32/float(sc.cameras[0].aspect)))
        plain=e.trace(p,32,hh,0,8,wavefront=True)
        L.emu_set_suspend_every(2)
        parked=e.trace(p,32,hh,0,8,wavefront=True)
        L.emu_set_suspend_every(0)
        assert np.array_equal(plain['image'], parked['image'])
        print('   resumed', L.emu_resumed_rays(1), 'stolen', L.emu_stolen_samples(1), flush=True)
    print('asan ok', s, flush=True)
# braided / flattened instance paths on small scenes (entry records at BLAS sub-trees, instance-space leaf tests)
for braid in ('1', '16'):
    os.environ['JT_BRAID_MAX'] = braid
    os.environ['JT_BRAID_MIN_INSTANCES'] = '1'
    for s in ['features1', 'classroom', 'synthetic_all']:
        sc = synth.make_scene(s) if s.startswith('synth') else orc.jt.load_scene(os.path.join(ROOT, 'assets', 'scenes', f'{s}.jtscene'))
        b=bvhm.make_scene_bvh(sc); Lt=lm.make_trace_lights(sc)
        o=orc.Oracle(sc,b,Lt); e=emu.Emu(sc,b,Lt)
        p=orc.make_params(resolution=48, samples=2, batch=2, sampler=1); w,h=o.make_state(p)
        rays=raygen.camera_rays(o,p,w,h,3000,seed=3); sec=raygen.secondary_rays(rays,o.intersect(rays),seed=4)
        for trav in (0,3): e.intersect(np.concatenate([rays,sec]),trav)
        p=orc.make_params(resolution=32, samples=2, batch=2, sampler=1, traversal=0)
        e.trace(p,32,max(1,int(32/float(sc.cameras[0].aspect))),0,2,wavefront=True)
        print('asan ok', s, 'braid', braid, e.stats()['flattened'], flush=True)
