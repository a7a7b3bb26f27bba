"""Independent gradient oracle: the SAME discrete schemes written as differentiable torch code, so that
torch.autograd through the unrolled steps gives the exact reverse-mode derivative a discrete adjoint must reproduce
(SURVEY.md section 8c "Gradient oracle").  Deliberately written without looking at oracle/petsc_ts.py's adjoint."""
import torch


def rk_unrolled(f, u0, schedule, A, b, c):
    """schedule: list of (t, h, out_slot).  Returns list of outputs (slot order) given u0 as slot 0."""
    outs = {0: u0}
    u = u0
    s = len(b)
    for (t, h, slot) in schedule:
        K = []
        for i in range(s):
            y = u
            for j in range(i):
                if A[i][j] != 0.0:
                    y = y + (h * A[i][j]) * K[j]
            K.append(f(t + c[i] * h, y))
        for j in range(s):
            if b[j] != 0.0:
                u = u + (h * b[j]) * K[j]
        if slot >= 0:
            outs[slot] = u
    return outs, u


def ark_unrolled_linear_im(f_im, f_ex, Jfun, u0, schedule, At, A, b, ct, c, batch_N=None):
    """IMEX with a LINEAR implicit part: stage solve (I/(h g) - J) exact, differentiable via torch.linalg.solve.
    Jfun() returns the dense Jacobian of f_im (possibly depending on trainable parameters)."""
    outs = {0: u0}
    u = u0
    s = len(b)
    for (t, h, slot) in schedule:
        Y, KI, KE = [], [], []
        for i in range(s):
            Z = u
            for j in range(i):
                if At[i][j] != 0.0:
                    Z = Z + (h * At[i][j]) * KI[j]
                if A[i][j] != 0.0:
                    Z = Z + (h * A[i][j]) * KE[j]
            if At[i][i] == 0.0:
                y = Z
                ki = f_im(t + ct[i] * h, y)
            else:
                shift = 1.0 / (h * At[i][i])
                J = Jfun()
                N = J.shape[0]
                M = shift * torch.eye(N, dtype=J.dtype) - J
                # f_im(y) = y @ J^T (per sample): shift (y - Z) - J y = 0  =>  y = shift M^{-1} Z
                y = torch.linalg.solve(M, (shift * Z).reshape(-1, N).T).T.reshape(Z.shape)
                ki = shift * (y - Z)
            Y.append(y)
            KI.append(ki)
            KE.append(f_ex(t + c[i] * h, y))
        for j in range(s):
            u = u + (h * b[j]) * (KI[j] + KE[j])
        if slot >= 0:
            outs[slot] = u
    return outs, u


def theta_unrolled(f, u0, schedule, theta, newton_iters=12):
    outs = {0: u0}
    u = u0
    for (t, h, slot) in schedule:
        rhs = u + (h * (1 - theta)) * f(t, u) if theta < 1 else u
        x = u
        n = u.numel()
        for _ in range(newton_iters):
            F = x - (h * theta) * f(t + h, x) - rhs
            J = torch.autograd.functional.jacobian(lambda v: f(t + h, v), x, create_graph=True).reshape(n, n)
            M = torch.eye(n, dtype=u.dtype) - (h * theta) * J
            x = x - torch.linalg.solve(M, F.reshape(-1)).reshape(x.shape)
        u = x
        if slot >= 0:
            outs[slot] = u
    return outs, u
