"""The mechanical Fortran -> C translator (oracle/refgen/f2c.py) on small programs whose results can be worked out by
hand: the pin of the oracle (tests/test_cpu_refpin.py) is only as good as the translator, so its rules are tested on their
own -- column-major arrays with declared lower bounds, DO trip counts (negative and zero-trip loops, bounds evaluated once),
statement functions, implicit typing, COMMON storage shared between units, ENTRY, GOTO, DO WHILE, array sections as
actual arguments (copy-in / copy-out), whole-array and section assignments, WHERE, integer division, ** with integer and
real exponents, the numeric intrinsics, expression association order.  No Fortran compiler exists here; the expected
values are computed in Python by the Fortran rules.  TEST INFRASTRUCTURE (needs gcc only)."""
import ctypes
import math
import os
import shutil
import subprocess
import sys
import tempfile

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle", "refgen"))
import f2c  # noqa: E402


def build(src, units):
    tr = f2c.Translator()
    tr.known_units |= set(units)
    tr.load(f2c.read_fixed_form(src), only=set(units))
    csrc = tr.emit()
    base = os.path.join(ROOT, "oracle", "_ref", "f2c_tests")          # git-ignored build area of the translated code
    os.makedirs(base, exist_ok=True)
    d = tempfile.mkdtemp(prefix="t_", dir=base)
    cp, so = os.path.join(d, "t.c"), os.path.join(d, "t.so")
    open(cp, "w").write(csrc)
    r = subprocess.run(["gcc", "-O2", "-fno-fast-math", "-ffp-contract=off", "-fPIC", "-std=gnu11", "-w", "-shared", "-o", so, cp, "-lm"],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[:3000] + "\n" + csrc[-3000:]
    lib = ctypes.CDLL(so)
    shutil.rmtree(d, ignore_errors=True)                              # the mapping stays valid
    return lib, csrc


def dptr(a):
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_double))


def iref(v):
    return ctypes.byref(ctypes.c_int(v))


def test_arrays_lower_bounds_and_loops():
    src = """
      subroutine fill (a, n, m)
      implicit none
      integer n, m, i, j, cnt
      real a(0:n,-1:m)
      cnt = 0
      do j=-1,m
        do i=0,n
          a(i,j) = 100.*j + i
        enddo
      enddo
c     negative step, zero-trip loop, bounds evaluated once
      do i=n,0,-2
        a(i,-1) = a(i,-1) + 0.5
      enddo
      do i=3,2
        a(0,0) = -999.
      enddo
      j = 2
      do i=1,j
        j = j + 5
        cnt = cnt + 1
      enddo
      a(n,m) = cnt
      return
      end
"""
    L, _ = build(src, ["fill"])
    n, m = 4, 2
    a = np.zeros((m + 2, n + 1))                       # Fortran a(0:n,-1:m) column-major == C [m+2][n+1]
    L.fill_(dptr(a), iref(n), iref(m))
    exp = np.array([[100.0 * j + i for i in range(n + 1)] for j in range(-1, m + 1)])
    exp[0, [4, 2, 0]] += 0.5
    exp[-1, -1] = 2.0                                   # the trip count was fixed at 2 before j changed
    assert np.array_equal(a, exp)


def test_statement_function_implicit_typing_and_association():
    src = """
      subroutine sf (x, y, r, k)
      real x, y, r(6)
      real avg, a, b
      integer k
      avg(a, b) = 0.5*(a + b)
      iq = 7/2
      r(1) = avg(x, y)*iq
      r(2) = x - y - 1.0 + 2.0**3**2
      r(3) = x/y*y
      r(4) = -x**2
      r(5) = x**k + y**(-k)
      r(6) = x**0.5 + 2**k
      return
      end
"""
    L, csrc = build(src, ["sf"])
    x, y, k = 3.7, 1.3, 3
    r = np.zeros(6)
    L.sf_(ctypes.byref(ctypes.c_double(x)), ctypes.byref(ctypes.c_double(y)), dptr(r), iref(k))
    exp = [0.5 * (x + y) * 3, ((x - y) - 1.0) + 2.0 ** 9, (x / y) * y, -(x * x), (x * x) * x + 1.0 / ((y * y) * y), math.pow(x, 0.5) + 8]
    assert r[0] == exp[0] and r[1] == exp[1] and r[2] == exp[2] and r[3] == exp[3] and r[5] == exp[5]
    assert abs(r[4] - exp[4]) <= 2e-16 * abs(exp[4])   # x**3: repeated multiplication, either association
    assert "pow(" in csrc


def test_common_entry_goto_dowhile():
    src = """
      subroutine acc (x)
      implicit none
      real x, total
      integer ncall
      common /store/ total, ncall
      total = total + x
      ncall = ncall + 1
      return
      entry accreset
      total = 0.
      ncall = 0
      return
      end

      subroutine report (t, n, steps)
      implicit none
      real t, total, v
      integer n, ncall, steps
      common /store/ total, ncall
      t = total
      n = ncall
c     collatz steps of n with goto; halving with do while
      steps = 0
      v = 27.
 10   continue
      if (v .eq. 1.) goto 20
      if (mod(int(v),2) .eq. 0) then
        v = v/2.
      else
        v = 3.*v + 1.
      endif
      steps = steps + 1
      goto 10
 20   continue
      do while (t .gt. 1.)
        t = t/2.
      enddo
      return
      end
"""
    L, _ = build(src, ["acc", "report"])
    L.accreset_()
    for v in (1.5, 2.5, 8.0):
        L.acc_(ctypes.byref(ctypes.c_double(v)))
    t, n, s = ctypes.c_double(0), ctypes.c_int(0), ctypes.c_int(0)
    L.report_(ctypes.byref(t), ctypes.byref(n), ctypes.byref(s))
    assert n.value == 3 and s.value == 111 and t.value == 12.0 / 16.0
    L.accreset_()
    L.report_(ctypes.byref(t), ctypes.byref(n), ctypes.byref(s))
    assert n.value == 0 and t.value == 0.0


def test_sections_where_and_intrinsics():
    src = """
      subroutine scale2 (v, n)
      implicit none
      integer n, i
      real v(n)
      do i=1,n
        v(i) = 2.*v(i)
      enddo
      return
      end

      subroutine driver (a, r)
      implicit none
      integer i
      real a(4,3), r(8), w(4)
      a(:,:) = 1.
      a(2:3,2) = 5.
      call scale2 (a(1,3), 4)
      call scale2 (a(2:3,2), 2)
      w(:) = a(:,2)
      where (w(:) .gt. 5.) w(:) = -w(:)
      r(1) = w(1) + w(2) + w(3) + w(4)
      r(2) = sign(3., -0.5) + abs(-2.5) + max(1., 4., 2.) + min(3, 7)
      r(3) = nint(2.5) + nint(-2.5) + int(-1.7) + mod(-7, 3)
      r(4) = sqrt(16.) + exp(0.) + log(1.) + alog(1.) + tanh(0.)
      r(5) = float(7/2) + real(2)/4
      r(6) = a(1,3) + a(4,3)
      r(7) = max(2, 3)/2 + 0.25
      r(8) = 0.
      do i=1,4
        if (a(i,2) .ge. 10. .and. .not. (i .eq. 3)) r(8) = r(8) + i
      enddo
      return
      end
"""
    L, _ = build(src, ["scale2", "driver"])
    a, r = np.zeros((3, 4)), np.zeros(8)
    L.driver_(dptr(a), dptr(r))
    assert np.array_equal(a[2], [2, 2, 2, 2]) and np.array_equal(a[1], [1, 10, 10, 1]) and np.array_equal(a[0], [1, 1, 1, 1])
    assert r[0] == 1 - 10 - 10 + 1
    assert r[1] == -3.0 + 2.5 + 4.0 + 3
    assert r[2] == 3 + (-3) + (-1) + (-1)          # nint rounds half away from zero, int truncates, mod keeps the sign of the dividend
    assert r[3] == 4.0 + 1.0
    assert r[4] == 3.0 + 0.5
    assert r[5] == 4.0 and r[6] == 1.25 and r[7] == 2.0
