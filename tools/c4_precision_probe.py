"""CPU experiment behind DESIGN.md's reading of the full-size config-4 parity number: the float64 oracle against the SAME
oracle with the GPU path's precision recipe — residuals / Jacobians evaluated in float32 from a float32 state, J^T W J, the
solve and the retraction in float64 — for 3 LM steps of the full BASELINE config 4.  If the two CPU runs differ by about as
much as the GPU differs from the float64 fixture (2.4e-5 relative after step 3), that gap is the float32 linearisation of a
far-from-converged iterate, not the elimination.   (~10 minutes of CPU.)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from islam_b200 import synth
from oracle import pvgo_oracle as po

f32 = np.float32


class MixedLM(po.SparseLM):
    def _res(self):
        c = lambda a: a.astype(f32)
        res = po.residuals(c(self.nodes), c(self.vels), self.edges, c(self.poses), c(self.drots), c(self.dtrans), c(self.dvels),
                           c(self.dts))
        return tuple(r.astype(np.float64) for r in res)

    def assemble(self, res):
        keep = self.nodes, self.vels, self.poses, self.drots
        orig = po.jacobian_blocks

        def jb(nodes, vels, edges, poses, drots, r0, r2):
            Jv, Jr = orig(nodes.astype(f32), vels.astype(f32), edges, poses.astype(f32), drots.astype(f32), r0.astype(f32),
                          r2.astype(f32))
            return Jv.astype(np.float64), Jr.astype(np.float64)
        po.jacobian_blocks = jb
        try:
            return super().assemble(res)
        finally:
            po.jacobian_blocks = orig
            self.nodes, self.vels, self.poses, self.drots = keep

    def _update(self, dn, dv, sign=1.0):
        super()._update(dn, dv, sign)
        self.nodes = self.nodes.astype(f32).astype(np.float64)          # the state is stored in float32 (the reference's dtype)
        self.vels = self.vels.astype(f32).astype(np.float64)


if __name__ == '__main__':
    g = synth.config4()
    fx = np.load(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests', 'golden', 'c4_oracle_steps.npz'))
    stride, mid = int(fx['stride']), int(fx['mid'])
    lm = MixedLM(g, np.float64, solver='splu', rollback='backup')
    for k in range(mid):
        lm.step()
        print('step', k + 1, 'mixed-precision loss', lm.history[-1]['loss'], 'float64 oracle', float(fx['losses'][k]),
              'rel diff %.2e' % (abs(lm.history[-1]['loss'] - fx['losses'][k]) / fx['losses'][k]), flush=True)
    n, v = lm.aligned(g.init_nodes[0])
    print('after step', mid, 'mixed-precision CPU run vs float64 oracle fixture:', po.rel_pose_error(n[::stride], fx['nodes_mid']), flush=True)
