"""Host-side time-step controller: the decisions PETSc's TSSolve loop / TSAdapt take between steps.

Pure python floats -- no tensors, no device -- so the same object drives (a) the generic path, where it is fed the
weighted error norm the device kernel reduced, and (b) the fused path, where a fixed-step schedule is unrolled on the
host before a single launch.  Time is always float64 on the host (petsc_adjoint.py:811,822).

Reference behaviour reproduced:
  * [PETSc] TSAdaptChoose_Basic: accept iff enorm <= 1; h_next = h * clip(safety * enorm^(-1/order), 0.1, 10),
    safety 0.9 (x0.5 when the previous attempt of this step was rejected too), dt in [1e-20, 1e50], <= 10 rejections.
  * [PETSc] TS_EXACTFINALTIME_MATCHSTEP (petsc_adjoint.py:640): shorten/stretch the next step to land on the next output
    time (within 1% -> exactly; within 2x -> halve), remember the un-shortened step and restore it after the hit.
  * pnode's tspanPostStep (petsc_adjoint.py:518-532): per-step step-size lists, steps-per-output-interval counters with
    the 1e-5 / 1e-3 hit tolerance, and the "fails to step on all the specified points" check (867-868).
"""
import math

SAFETY = 0.9
REJECT_SAFETY = 0.5
CLIP_LO, CLIP_HI = 0.1, 10.0
DT_MIN, DT_MAX = 1e-20, 1e50
MATCH_NEAR, MATCH_HALF = 0.01, 2.0
MAX_REJECT_DEFAULT = 10
SPAN_RELTOL = 1e-6
SPAN_ABSTOL = 10 * 2.220446049250313e-16
SQRT_EPS = 1.4901161193847656e-08


def adapt_basic(h, enorm, order, prev_attempt_accepted):
    """Returns (accept, h_next) for one attempt; h_next is before the MATCHSTEP clamp."""
    safety = SAFETY
    if enorm > 1.0:
        if not prev_attempt_accepted:
            safety *= REJECT_SAFETY
        accept = h < (1.0 + SQRT_EPS) * DT_MIN
    else:
        accept = True
    fac = safety * enorm ** (-1.0 / order) if enorm > 0.0 else math.inf
    fac = min(max(fac, CLIP_LO), CLIP_HI)
    return accept, min(max(h * fac, DT_MIN), DT_MAX)


class TimeLoop:
    """State machine of one TSSolve.  Usage:
           while not loop.done:
               t, h = loop.t, loop.h            # attempt a step of size h from t
               accepted = loop.report(enorm)    # enorm = None when not adaptive
       After an accepted step `loop.last_out_slot` is the span slot u_{n+1} must be copied to (or -1)."""

    def __init__(self, times, step_size, adaptive, order, double_precision, max_reject=MAX_REJECT_DEFAULT):
        times = [float(x) for x in times]
        if len(times) == 1:  # integrate [0, t0], no span bookkeeping (petsc_adjoint.py:818-820)
            self.span = None
            self.t = 0.0
            self.t_end = times[0]
        else:
            self.span = times
            self.t = times[0]
            self.t_end = times[-1]
        self.step_list = step_size if isinstance(step_size, list) else None
        self.h = float(step_size[0] if self.step_list is not None else step_size)
        self.adaptive = adaptive
        self.order = order
        self.max_reject = max_reject
        self.delta = 1e-5 if double_precision else 1e-3
        self.steps = 0
        self.ctr = 1  # next un-hit span slot ([PETSc] tspan->spanctr)
        self.cur_sol_steps = [0] * len(times)
        self.cur_sol_index = 1
        self._prev_ok = True
        self._rejections = 0
        self._dt_span_cached = 0.0
        self.last_out_slot = -1
        # [PETSc] TSSolve, before the first step: MATCHSTEP keeps the initial step from overshooting the first target
        first = self.span[1] if self.span is not None else self.t_end
        if self.h >= first - self.t or abs(self.h - (first - self.t)) <= SPAN_ABSTOL:
            self.h = first - self.t
        self.last_h = self.h
        self.attempts = []  # (t, h, accepted, enorm)

    @property
    def done(self):
        return not (self.t < self.t_end and abs(self.t - self.t_end) > SPAN_ABSTOL)

    # [PETSc] TSAdaptChoose tail: MATCHSTEP clamp, evaluated at the time reached by the step just accepted
    def _matchstep(self, t_new, h, h_next):
        hit = False
        if self.span is not None:
            c = min(self.ctr, len(self.span) - 1)
            if abs(t_new - self.span[c]) <= SPAN_RELTOL * h + SPAN_ABSTOL:
                hit = True
                tmax = self.span[c + 1] if c + 1 < len(self.span) else self.t_end
            else:
                tmax = self.span[c]
        else:
            tmax = self.t_end
        out = h_next
        tend = t_new + h_next
        hmax = tmax - t_new
        if t_new < tmax:
            if tend > tmax:
                out = hmax
            elif tend < tmax:
                if h_next * MATCH_HALF > hmax:
                    out = hmax / 2
                if h_next * (1.0 + MATCH_NEAR) > hmax:
                    out = hmax
        if self.span is not None:
            if h != out and not self._dt_span_cached:
                self._dt_span_cached = h
            if h == out and self._dt_span_cached and hit:
                out = self._dt_span_cached
                self._dt_span_cached = 0.0
        return out

    def report(self, enorm=None):
        h = self.h
        if self.adaptive:
            accept, h_next = adapt_basic(h, enorm, self.order, self._prev_ok)
        else:
            accept, h_next = True, h
        self.attempts.append((self.t, h, accept, -1.0 if enorm is None else enorm))
        if not accept:
            self.h = h_next
            self._prev_ok = False
            self._rejections += 1
            if self._rejections > self.max_reject:
                raise RuntimeError("TS_DIVERGED_STEP_REJECTED: %d consecutive rejections at t=%g" %
                                   (self._rejections, self.t))
            return False
        t_new = self.t + h
        h_next = self._matchstep(t_new, h, h_next)
        self._prev_ok = True
        self._rejections = 0
        self.t = t_new
        self.steps += 1
        self.last_h = h
        self.h = h_next
        # pnode's PostStep hook (only installed for span runs)
        if self.span is not None and self.cur_sol_index < len(self.span):
            if self.step_list is not None and self.steps < len(self.step_list):
                self.h = float(self.step_list[self.steps])
            self.cur_sol_steps[self.cur_sol_index] += 1
            if abs(t_new - self.span[self.cur_sol_index]) < self.delta:
                self.cur_sol_index += 1
        # [PETSc] TSSolve: copy the solution into the span slot when the step landed on it
        self.last_out_slot = -1
        if self.span is not None and self.ctr < len(self.span) and \
                abs(t_new - self.span[self.ctr]) <= SPAN_RELTOL * h + SPAN_ABSTOL:
            self.last_out_slot = self.ctr
            self.ctr += 1
        return True

    def follow(self, accept, enorm, t_next, h_next):
        """Book-keeping of one attempt whose verdict the DEVICE controller took (csrc/cnf_rk.cu, namespace ctl): the
        attempt started at (self.t, self.h); `accept` and the next attempt's (t_next, h_next) come from the device's log,
        so the host never re-derives a step size the kernels did not use.  Mirrors report()."""
        h = self.h
        self.attempts.append((self.t, h, bool(accept), enorm))
        if not accept:
            self.h = h_next
            self._prev_ok = False
            self._rejections += 1
            return False
        t_new = t_next
        self._prev_ok = True
        self._rejections = 0
        self.t = t_new
        self.steps += 1
        self.last_h = h
        self.h = h_next
        if self.span is not None and self.cur_sol_index < len(self.span):
            self.cur_sol_steps[self.cur_sol_index] += 1
            if abs(t_new - self.span[self.cur_sol_index]) < self.delta:
                self.cur_sol_index += 1
        self.last_out_slot = -1
        if self.span is not None and self.ctr < len(self.span) and \
                abs(t_new - self.span[self.ctr]) <= SPAN_RELTOL * h + SPAN_ABSTOL:
            self.last_out_slot = self.ctr
            self.ctr += 1
        return True

    def check_complete(self):
        if self.span is not None and self.cur_sol_index != len(self.span):
            raise Exception("TSSolve fails to step on all the specified points")  # petsc_adjoint.py:867-868


def fixed_schedule(times, step_size, double_precision):
    """Unroll a non-adaptive run on the host.  Returns (loop, steps) with steps = [(t_n, h_n, out_slot)]."""
    loop = TimeLoop(times, step_size, adaptive=False, order=1, double_precision=double_precision)
    steps = []
    while not loop.done:
        t, h = loop.t, loop.h
        loop.report(None)
        steps.append((t, h, loop.last_out_slot))
        if len(steps) > 50_000_000:
            raise RuntimeError("fixed_schedule: step size too small for the requested interval")
    return loop, steps
