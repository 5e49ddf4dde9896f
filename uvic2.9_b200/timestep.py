"""Host-side time stepping of the tracer path, as `mom` sequences it.

source/mom/mom.F:85-146 and 09/common/switch.F:217-224: with nmix > 1 every step whose
counter satisfies mod(itt, nmix) == 1 is a forward "mixing" step (tau-1 := tau, c2dtts = dtts);
all other steps are leapfrog (c2dtts = 2 dtts).  After each step the time levels rotate
(source/mom/mom.F:210-212).  run/control.in sets nmix=16, eb=.false.
"""
from __future__ import annotations


def is_leapfrog(itt: int, nmix: int = 16) -> bool:
    if nmix in (0, 1):
        return True
    return itt % nmix != 1


class TimeStepper:
    def __init__(self, ctx, nmix=16, itt0=0):
        self.ctx = ctx
        self.nmix = nmix
        self.itt = itt0

    def advance(self, nsteps=1, diag=False):
        for _ in range(nsteps):
            self.itt += 1
            self.ctx.step(leapfrog=is_leapfrog(self.itt, self.nmix), diag=diag)
            self.ctx.rotate()
