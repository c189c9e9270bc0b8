"""Role-level cycle breakdown of the tensor-core kernel (CTA 0): render-core launch and SDF-only launches."""
import ctypes as C, os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import __graft_entry__ as g; g.build()
import bench, color_neus_b200 as cn
from color_neus_b200 import _lib as L
from color_neus_b200.rays import synthetic_camera_rays
lib = L.lib()
lib.cneus_tc_prof_enable.argtypes = [C.c_int]; lib.cneus_tc_prof_read.argtypes = [C.POINTER(C.c_ulonglong), C.c_int]
torch.manual_seed(1)
ren = cn.Color_NeuS(bench.renderer_cfg()).cuda().eval()
ro, rd, near, far = synthetic_camera_rays(800, 800, device="cuda")
n = int(sys.argv[1]) if len(sys.argv) > 1 else 32768
PW = int(os.environ.get('CNEUS_PROF_WARP', '0'))   # epilogue warp whose timeline is recorded; 99: role counters only (no timeline marks)
sl = slice(300 * 800, 300 * 800 + n)
names = ["mma_wait_A", "mma_wait_W", "mma_total", "steps", "epi_wait_acc", "epi_total", "prod_wait_slot", "prod_total",
         "wait_slab0", "wait_slab1", "wait_slab2", "wait_slab3", "w12_wait_acc", "w12_total", "-", "-",
         "e_sec01", "e_bar1", "e_sig01", "e_sec2", "e_sig2", "e_sec3", "e_skipfeed", "e_park", "e_bwdlast", "e_sig3", "e_post",
         "e_sig_all", "e_tail+bias", "e_seed", "e_color_in", "e_relight_in", "e_cg"]
with torch.no_grad():
    for _ in range(2): ren(ro[sl], rd[sl], near[sl], far[sl])
    z = ren._last["z_vals"]
    torch.cuda.synchronize()
    for label, fn in (("render_core", lambda: ren.render_core(ro[sl], rd[sl], z, 2.0 / 64)), ("sample_z (4 sdf-only launches)", lambda: ren.sample_z(ro[sl], rd[sl], near[sl], far[sl], None))):
        out = (C.c_ulonglong * 32)()
        lib.cneus_tc_prof_enable(1 + PW); lib.cneus_tc_prof_read(out, 1)
        if hasattr(lib, "cneus_tc_prof_read_types"): lib.cneus_tc_prof_read_types((C.c_ulonglong * 16)(), 1)
        fn(); torch.cuda.synchronize()
        lib.cneus_tc_prof_read(out, 1); lib.cneus_tc_prof_enable(0)
        v = list(out); steps = max(v[3], 1)
        print(label, {k: int(x) for k, x in zip(names, v)})
        print("  per step: mma_total %.0f (wait_A %.0f, wait_W %.0f, issue+drain %.0f) | epi wait_acc %.0f work %.0f | producer wait %.0f" % (
            v[2] / steps, v[0] / steps, v[1] / steps, (v[2] - v[0] - v[1]) / steps, v[4] / steps, (v[5] - v[4]) / steps, v[6] / steps))
        print("  per step: MMA waits for slab 0..3: %.0f %.0f %.0f %.0f | warp 12: wait_acc %.0f work %.0f" % (
            v[8] / steps, v[9] / steps, v[10] / steps, v[11] / steps, v[12] / steps, (v[13] - v[12]) / steps))
        if hasattr(lib, "cneus_tc_prof_read_types"):
            ty = (C.c_ulonglong * 16)(); lib.cneus_tc_prof_read_types(ty, 1); ty = list(ty)
            tn = ["softplus+save", "softplus", "grad chain", "relu", "feature block", "encoding adjoint"]
            if sum(ty):
                print("  epilogue cycles per step by type:", " | ".join("%s %.0f (x%d)" % (tn[i], ty[2 * i] / max(ty[2 * i + 1], 1), ty[2 * i + 1]) for i in range(6) if ty[2 * i + 1]))
        if sum(v[16:32]):
            print("  epilogue timeline per step:", " ".join("%s %.0f" % (k[2:], x / steps) for k, x in zip(names[16:], v[16:32]) if k != "-"))
