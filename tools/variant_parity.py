import sys
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests')
import parity, dualsphysics_multilayer_b200 as dsph
from dualsphysics_multilayer_b200 import capi
L=dsph.lib()
for name in ('pb_sym_wend_art_ddt1','db_sym_wend_art_ddt2','db2d_sym_wend_art_ddt2_shift'):
    g=parity.Golden(name)
    for variant in (1,4,6,8):
        ctx=parity.make_ctx(L,g.cfg,20000)
        L.lib.dsph_set_force_variant(ctx.h,variant)
        step=g.steps[-1]
        s0=g.stage(step,'div0'); parity.upload_state(ctx,s0,g.cfg); ctx.celldiv(False)
        fr=ctx.forces(2); ace,ar=ctx.forces_download(); ref=g.stage(step,'fpre'); npb=int(ref['npb'])
        print(name,'variant',variant,'ace %.2e ar %.2e'%(parity.rl2(ace[npb:],ref['ace'][npb:]),parity.rl2(ar,ref['ar'])))
        ctx.close()
