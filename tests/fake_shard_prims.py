"""CPU test double for the ``pdlp_shard_*`` C-ABI calls (TEST INFRASTRUCTURE ONLY).

Implements the per-rank device work of torchpdlp_b200.sharded with torch CPU tensors so that the
host-side orchestration (row partition, the sequence of collectives, deferred step rule, restart
protocol) can be exercised under gloo with world_size 2 on a machine without GPUs.  It follows the
same arithmetic as the CUDA kernels (pdlp_ops.cuh) and, through them, the reference lines cited there.
"""
import math
from types import SimpleNamespace

import numpy as np
import torch


class FakeShardPrims:
    def __init__(self, K_local, m_ineq_local, c, q_local, l, u):
        import scipy.sparse as sp
        self.K = sp.csr_matrix(K_local)
        self.Kt = sp.csr_matrix(self.K.T)
        t = lambda v: torch.as_tensor(np.asarray(v, dtype=np.float64)).reshape(-1).clone()
        self.c, self.q, self.l, self.u = t(c), t(q_local), t(l), t(u)
        self.m_local, self.n = self.K.shape
        self.m_ineq = int(m_ineq_local)
        self.device = torch.device("cpu")
        z = torch.zeros
        self.x, self.dx, self.y = z(self.n), z(self.n), z(self.m_local)
        self.xsum, self.ysum = z(self.n), z(self.m_local)
        self.xlast, self.ylast = z(self.n), z(self.m_local)
        self.xprev, self.yprev = z(self.n), z(self.m_local)
        self.s = SimpleNamespace(eta=1.0, omega=1.0, theta=1.0, w_pending=0.0, eta_sum=0.0, dx2=0.0, dy2=0.0, inter=0.0,
                                 last_weight=0.0, kkt_first=0.0, q_norm=0.0, c_norm=float(torch.linalg.norm(self.c)),
                                 tol=1e-4, k=0, t=0, n_restarts=0, adaptive=0, primal_update=0, rule_pending=0,
                                 infeas=0, infeas_k=0, solved=0, have_prev=False)

    # ---- plumbing ----
    def buffer(self, length):
        return torch.zeros(length, dtype=torch.float64)

    def enable(self, q_norm, c_norm):
        self.s.q_norm, self.s.c_norm, self.s.rule_pending = q_norm, c_norm, 0

    def local_q_sumsq(self):
        return float(torch.dot(self.q, self.q)), float(torch.linalg.norm(self.c))

    def spmv(self, v, transpose):
        A = self.Kt if transpose else self.K
        return torch.from_numpy(A @ v.numpy())

    def set_iterate(self, x, y_local):
        self.x = torch.zeros(self.n) if x is None else torch.as_tensor(x, dtype=torch.float64).reshape(-1).clone()
        self.y = torch.zeros(self.m_local) if y_local is None else torch.as_tensor(y_local, dtype=torch.float64).reshape(-1).clone()
        self.dx = torch.zeros(self.n)
        self.xsum.zero_(); self.ysum.zero_()
        self.xlast, self.ylast = self.x.clone(), self.y.clone()
        self.s.w_pending = self.s.eta_sum = 0.0
        self.s.t = 0

    def get_iterate(self):
        return self.x.clone(), self.y.clone()

    def set_step(self, eta, omega, theta, adaptive):
        self.s.eta, self.s.omega, self.s.theta, self.s.adaptive = eta, omega, theta, int(adaptive)

    def set_options(self, tol, primal_update):
        self.s.tol, self.s.primal_update = tol, int(primal_update)

    def set_counters(self, k, t):
        self.s.k, self.s.t = k, t

    def state(self):
        return self.s

    def sync(self):
        pass

    # ---- iteration ----
    def kty(self, buf):
        buf[:self.n] = self.spmv(self.y, True)
        buf[self.n] = self.s.dy2 if self.s.rule_pending else 0.0
        buf[self.n + 1] = self.s.inter if self.s.rule_pending else 0.0
        buf[self.n + 2:] = 0.0

    def _rule(self, dy2, inter):
        s = self.s
        if not s.rule_pending:
            return
        s.dy2, s.inter = dy2, inter
        s.k += 1
        k, eta, omega = s.k, s.eta, s.omega
        w = nxt = eta
        if s.adaptive:
            den = 2.0 * inter
            if den != 0.0:
                eta_bar = (omega * s.dx2 + dy2 / omega) / abs(den)
                t1 = (1.0 - (k + 1) ** (-0.3)) * eta_bar
            else:
                eta_bar = t1 = math.inf
            ep = min(t1, (1.0 + (k + 1) ** (-0.6)) * eta)
            w, nxt = (eta, ep) if eta <= eta_bar else (ep, ep)
        s.eta_sum += w
        s.w_pending = s.last_weight = w
        s.eta = nxt
        s.t += 1
        s.rule_pending = 0

    def step(self, buf, save_previous):
        s = self.s
        self._rule(float(buf[self.n]), float(buf[self.n + 1]))
        tau, sigma, wp = s.eta / s.omega, s.eta * s.omega, s.w_pending
        g = self.c - buf[:self.n]
        xo = self.x
        xn = torch.minimum(torch.maximum(xo - tau * g, self.l), self.u)
        self.dx = xn - xo
        self.xsum += wp * xo
        if save_previous:
            self.xprev, s.have_prev = xo.clone(), True
        self.x = xn
        s.dx2 = float(torch.dot(self.dx, self.dx))
        kxbar = self.spmv(xn + s.theta * self.dx, False)
        kdx = self.spmv(self.dx, False)
        yo = self.y
        yn = yo + sigma * (self.q - kxbar)
        yn[:self.m_ineq] = torch.clamp(yn[:self.m_ineq], min=0.0)
        dy = yn - yo
        self.ysum += wp * yo
        if save_previous:
            self.yprev = yo.clone()
        self.y = yn
        s.dy2, s.inter, s.rule_pending = float(torch.dot(dy, dy)), float(torch.dot(dy, kdx)), 1

    # ---- exchange="p2p": the same device work, the exchange done by the "kernels" (here: the process group) ----
    def exchange_connect(self, group=None):
        import torch.distributed as dist
        self._group = group
        self._world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.connected = True
        self.p2p_exchanges = 0

    def exchange_disconnect(self):
        self.connected = False

    def allreduce_p2p(self, t):
        import torch.distributed as dist
        assert self.connected
        if self._world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.SUM, group=self._group)
        self.p2p_exchanges += 1

    def iterate_p2p(self, n_iters, save_previous):
        buf = self.buffer(self.n + 4)
        for i in range(n_iters):
            self.kty(buf)
            self.allreduce_p2p(buf)
            self.step(buf, save_previous and i == n_iters - 1)

    def rule_pack(self, buf4):
        buf4[0] = self.s.dy2 if self.s.rule_pending else 0.0
        buf4[1] = self.s.inter if self.s.rule_pending else 0.0
        buf4[2:] = 0.0

    def rule_apply(self, buf4):
        self._rule(float(buf4[0]), float(buf4[1]))

    # ---- restart test ----
    def check_begin(self, buf3, have_prev):
        s = self.s
        s.have_prev = bool(have_prev)
        self.xsum += s.w_pending * self.x
        self.ysum += s.w_pending * self.y
        es = s.eta_sum
        self.X3 = [self.x, self.xsum / es if es > 0 else self.x, self.xprev if s.have_prev else self.x]
        self.Y3 = [self.y, self.ysum / es if es > 0 else self.y, self.yprev if s.have_prev else self.y]
        out = torch.stack([self.spmv(yv, True) for yv in self.Y3], dim=1)
        buf3.copy_(out.reshape(-1))

    def check_mid(self, buf3, buf9):
        kty3 = buf3.reshape(self.n, 3)
        lo, up = torch.isinf(self.l) & (self.l < 0), torch.isinf(self.u) & (self.u > 0)
        ld, ud = torch.where(lo, torch.zeros_like(self.l), self.l), torch.where(up, torch.zeros_like(self.u), self.u)
        self.cols = []
        for v in range(3):
            g = self.c - kty3[:, v]
            lam = torch.where(lo & up, torch.zeros_like(g), torch.where(lo, torch.clamp(g, max=0.0),
                              torch.where(up, torch.clamp(g, min=0.0), g)))
            d = g - lam
            self.cols.append((float(torch.dot(d, d)), float(torch.dot(self.c, self.X3[v])),
                              float(torch.dot(ld, torch.clamp(lam, min=0.0))), float(torch.dot(ud, torch.clamp(lam, max=0.0)))))
            r = self.spmv(self.X3[v], False) - self.q
            r[:self.m_ineq] = torch.clamp(r[:self.m_ineq], max=0.0)
            buf9[3 * v] = float(torch.dot(r, r))
            buf9[3 * v + 1] = float(torch.dot(self.q, self.Y3[v]))
            buf9[3 * v + 2] = buf9[3 * v]
        self.s.w_pending = 0.0

    def check_decide(self, buf9, force, buf1):
        s = self.s
        res = []
        for v in range(3):
            dr2, ctx, lpos, uneg = self.cols[v]
            pr, dr = math.sqrt(float(buf9[3 * v])), math.sqrt(dr2)
            adj = float(buf9[3 * v + 1]) + lpos + uneg
            gap = adj - ctx
            w2 = s.omega ** 2
            res.append(dict(pr=pr, dr=dr, gap=gap, prim_obj=ctx, adj_dual=adj,
                            kkt=math.sqrt(w2 * pr * pr + dr * dr / w2 + gap * gap)))
        kc, ka, kp = res[0]["kkt"], res[1]["kkt"], res[2]["kkt"]
        kmin, use_avg, kind = min(kc, ka), int(kc >= ka), 0
        if force:
            kind, use_avg = 4, 0
        elif kmin <= 0.2 * s.kkt_first:
            kind = 1
        elif kmin <= 0.8 * s.kkt_first and kmin > kp:
            kind = 2
        elif s.t >= 0.36 * s.k:
            kind = 3
        self.chk = SimpleNamespace(kind=kind, use_average=use_avg, t=s.t, k=s.k, solved=0, kkt_current=kc, kkt_average=ka,
                                   kkt_previous=kp, point=SimpleNamespace(**res[1 if use_avg else 0], pr_unscaled=0.0, dr_unscaled=0.0))
        buf1[0] = 0.0
        self.dxr2 = 0.0
        if kind:
            xc, yc = self.X3[1 if use_avg else 0].clone(), self.Y3[1 if use_avg else 0].clone()
            self.dxr2 = float(torch.dot(self.xlast - xc, self.xlast - xc))
            buf1[0] = float(torch.dot(self.ylast - yc, self.ylast - yc))
            self.x, self.y, self.dx = xc, yc, torch.zeros(self.n)
            self.xlast, self.ylast = xc.clone(), yc.clone()
            self.xsum.zero_(); self.ysum.zero_()

    def check_end(self, buf1):
        s, chk = self.s, self.chk
        s.have_prev = False
        if chk.kind:
            s.n_restarts += 1
            if s.primal_update:
                dxn, dyn = math.sqrt(self.dxr2), math.sqrt(float(buf1[0]))
                if dxn > 0 and dyn > 0:
                    s.omega = math.exp(0.5 * math.log(dyn / dxn) + 0.5 * math.log(s.omega))
            p = chk.point
            w2 = s.omega ** 2
            p.kkt = math.sqrt(w2 * p.pr ** 2 + p.dr ** 2 / w2 + p.gap ** 2)
            s.kkt_first = p.kkt
            chk.solved = int(p.pr <= s.tol * (1 + s.q_norm) and p.dr <= s.tol * (1 + s.c_norm)
                             and p.gap <= s.tol * (1 + abs(p.prim_obj) + abs(p.adj_dual)))
            s.t, s.eta_sum, s.w_pending = 0, 0.0, 0.0
        chk.omega = s.omega
        return chk
