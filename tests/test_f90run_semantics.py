"""oracle/f90run on a Fortran module written for the purpose: the language rules the reference-run parity tests lean on, each with a
value that follows from the Fortran standard.  (That the translator reproduces the REFERENCE is checked by running the reference's
own unit tests through it, tests/test_reference_unit_tests.py; this file documents the constructs one by one and needs no reference tree.)"""
import pytest

from oracle import f90run
from oracle.f90run.rt import FArray

SOURCE = r'''
module sem
implicit none ; private
public :: run_all, shape2d, circle, square, area_of, half

type, abstract :: shape2d
  real :: scale = 1.0
contains
  procedure(area_if), deferred :: area
end type shape2d
abstract interface
  real function area_if(this)
    import shape2d
    class(shape2d), intent(in) :: this
  end function area_if
end interface
type, extends(shape2d) :: circle
  real :: r = 0.0
contains
  procedure :: area => circle_area
end type circle
type, extends(shape2d) :: square
  real :: a = 0.0
contains
  procedure :: area => square_area
end type square

type :: pair
  real :: x(3)
  integer :: n = 0
end type pair

interface split
  module procedure split_scalar, split_array
end interface split

contains

real function circle_area(this)
  class(circle), intent(in) :: this
  circle_area = this%scale * (3.0 * this%r**2)
end function circle_area

real function square_area(this)
  class(square), intent(in) :: this
  square_area = this%scale * this%a**2
end function square_area

!> which arm of SELECT TYPE runs, and a type-bound call on the associate name
subroutine area_of(s, arm, a)
  class(shape2d), intent(in) :: s
  integer, intent(out) :: arm
  real, intent(out) :: a
  select type (t => s)
    type is (circle)
      arm = 1 ; a = t%area()
    type is (square)
      arm = 2 ; a = t%area()
    class default
      arm = 3 ; a = -1.0
  end select
end subroutine area_of

!> a function that defines an argument: the translator returns function values only and must say so rather than lose the flag
real function half(x, odd)
  integer, intent(in) :: x
  logical, intent(out) :: odd
  odd = (mod(x, 2) /= 0)
  half = 0.5 * real(x)
end function half

subroutine split_scalar(x, whole, frac)
  real, intent(in) :: x
  integer, intent(out) :: whole
  real, intent(out) :: frac
  whole = int(x) ; frac = x - real(whole)
end subroutine split_scalar

subroutine split_array(x, whole, frac)
  real, intent(in) :: x(:)
  integer, intent(out) :: whole
  real, intent(out) :: frac
  whole = int(sum(x)) ; frac = sum(x) - real(whole)
end subroutine split_array

integer function counter()
  integer, save :: calls = 0
  calls = calls + 1
  counter = calls
end function counter

real elemental function twice(x)
  real, intent(in) :: x
  twice = 2.0 * x
end function twice

subroutine shifted(a, lo, total)
  integer, intent(in) :: lo
  real, dimension(lo:), intent(in) :: a
  real, intent(out) :: total
  integer :: i
  total = 0.0
  do i = lo, ubound(a, 1) ; total = total + real(i) * a(i) ; enddo
end subroutine shifted

!> everything returned through res(:) so that one call checks the lot
subroutine run_all(res, text)
  real, intent(out) :: res(:)
  character(len=*), intent(out) :: text
  real :: a(5), f, tot, host_var
  integer :: w, i, k
  type(pair) :: p, q
  type(circle) :: c
  type(square) :: s
  res(:) = 0.0
  ! integer division truncates toward zero; real ** integer by repeated multiplication; mixed arithmetic
  res(1) = real((-7) / 2) ; res(2) = real(7 / 2 * 2) ; res(3) = 2.0**(-2) ; res(4) = real(2**10) ; res(5) = 7 / 2.0
  ! array sections, lower bounds re-based on argument association
  do i = 1, 5 ; a(i) = real(i) ; enddo
  call shifted(a(2:4), -1, tot) ; res(6) = tot                    ! (-1)*2 + 0*3 + 1*4
  ! generic subroutines define scalar arguments whichever specific runs
  call split(3.75, w, f) ; res(7) = real(w) ; res(8) = f
  call split(a(1:3), w, frac=f) ; res(9) = real(w) ; res(10) = f
  ! SAVE
  k = counter() ; k = counter() ; res(11) = real(counter())
  ! elemental on an array section; MAX / MIN / SIGN / MOD / NINT
  a(1:2) = twice(a(4:5)) ; res(12) = a(1) + a(2)
  res(13) = max(1.0, 3.0, 2.0) + min(4, 2) + sign(2.0, -0.0) + mod(-7, 3) + nint(2.5) + nint(-2.5)
  ! derived-type assignment copies (value semantics)
  p%x(:) = 1.0 ; p%n = 4 ; q = p ; q%x(2) = 9.0 ; q%n = 5 ; res(14) = p%x(2) + real(p%n)
  ! internal procedure with host association (reads and defines a host variable)
  host_var = 10.0 ; call bump(2.5) ; res(15) = host_var
  ! select type and type-bound procedures
  c%r = 2.0 ; c%scale = 0.5 ; s%a = 3.0
  call area_of(c, k, f) ; res(16) = real(k) + f ; call area_of(s, k, f) ; res(17) = real(k) + f
  ! do loop variable after the loop, exit / cycle, do while
  do i = 1, 10 ; if (i == 3) cycle ; if (i > 6) exit ; enddo ; res(18) = real(i)
  i = 0 ; do while (i < 4) ; i = i + 3 ; enddo ; res(19) = real(i)
  ! select case with ranges
  select case (nint(res(18)))
    case (:3) ; res(20) = 1.0
    case (5, 7:9) ; res(20) = 2.0
    case default ; res(20) = 3.0
  end select
  write(text, '(I3,",",F7.3,",",ES10.3)') w, f, tot
contains
  subroutine bump(by)
    real, intent(in) :: by
    host_var = host_var + by * real(w)
  end subroutine bump
end subroutine run_all

end module sem
'''


@pytest.fixture(scope="module")
def sem(tmp_path_factory):
    p = tmp_path_factory.mktemp("f90") / "sem.F90"
    p.write_text(SOURCE)
    return f90run.load([str(p)])["sem"]


def test_language_rules(sem):
    res = FArray.alloc("r", [(1, 20)])
    (text,) = sem["run_all"](res, "")
    want = [-3.0,    # (-7)/2: integer division truncates toward zero
            6.0,     # 7/2*2, left to right in integers
            0.25, 1024.0, 3.5,
            2.0,     # a(2:4) seen as a(-1:1) by the callee: (-1)*2 + 0*3 + 1*4
            3.0, 0.75,   # generic subroutine, scalar specific: both intent(out) scalars come back
            6.0, 0.0,    # ... array specific, one of them by keyword
            3.0,     # SAVE: the third call
            18.0,    # elemental function over a section
            2.0,     # max(1,3,2) + min(4,2) + sign(2,-0.0) + mod(-7,3) + nint(2.5) + nint(-2.5) = 3 + 2 - 2 - 1 + 3 - 3
            5.0,     # q = p copies: changing q leaves p%x(2) = 1 and p%n = 4
            25.0,    # the internal procedure reads w (= 6) and defines host_var: 10 + 2.5 * 6
            7.0,     # select type: circle, arm 1, area 0.5 * 3 * 2**2
            11.0,    # square, arm 2, area 9
            7.0,     # the do variable after EXIT at i = 7
            6.0,     # do while
            2.0]     # select case (7) falls in 7:9
    assert res.tolist() == want
    assert text == "  6,  9.000, 2.000E+00"   # write(text, (I3,",",F7.3,",",ES10.3)) w, f, tot


def test_a_function_that_defines_an_argument_stops_the_run(sem):
    with pytest.raises(NotImplementedError, match="changed its scalar argument odd"):
        sem["half"](3, False)
    assert sem["half"](4, False) == 2.0   # nothing lost: the flag keeps the value it came in with
