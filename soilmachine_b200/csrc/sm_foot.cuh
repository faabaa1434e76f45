// sm_foot.cuh -- the cell sets the exact-footprint schedules order steps by, as closed-form predicates.
// Pure integer geometry, host + device: tests/hostsim checks every predicate against explicit cell sets
// (tests/test_host_build.py::test_footprint_predicates_equal_the_cell_sets).
//   water: move() reads plus(ip); the step touches plus(ip) U 3x3(np), writes {ip} U 3x3(np)   (no nested re-cascade)
//   wind : move() reads plus(ip); the step touches and writes 5x5(ip) U 5x5(np)                (cascade(.,1) + one re-cascade)
// ip = ipos, np = npos (known after move), R = published reach (box ip +- R holds the whole step).  d* = B - A.
#pragma once
#include "sm_core.cuh"

SM_HD bool plus_hits_3x3(int dx, int dy) {   // plus(c) meets 3x3(c + d)
  dx = dx < 0 ? -dx : dx; dy = dy < 0 ? -dy : dy;
  return (dx <= 1 && dy <= 2) || (dx <= 2 && dy <= 1);
}
SM_HD bool plus_hits_box3(int dx, int dy) {  // plus(c) meets the box (c + d) +- 3
  dx = dx < 0 ? -dx : dx; dy = dy < 0 ? -dy : dy;
  return (dx <= 4 && dy <= 3) || (dx <= 3 && dy <= 4);
}
SM_HD int iabs_(int v) { return v < 0 ? -v : v; }
SM_HD bool plus_hits_5x5(int dx, int dy) {   // plus(c) meets 5x5(c + d)
  dx = iabs_(dx); dy = iabs_(dy);
  return (dx <= 3 && dy <= 2) || (dx <= 2 && dy <= 3);
}
template <int KIND> struct Foot;
template <> struct Foot<0> {
  static SM_HD bool in_range(int dx, int dy, int, int) { return iabs_(dx) <= 6 && iabs_(dy) <= 6; }
  // d* = B - A
  static SM_HD bool box_hits_M(int dx, int dy, int) { return plus_hits_box3(dx, dy); }
  static SM_HD bool W_hits_M(int ibx, int iby, int nbx, int nby, int ax, int ay) {
    return (iabs_(ibx - ax) + iabs_(iby - ay) <= 1) || plus_hits_3x3(nbx - ax, nby - ay);
  }
  static SM_HD bool F_hits_F(int ax, int ay, int nax, int nay, int bx, int by, int nbx, int nby) {
    return (iabs_(bx - ax) + iabs_(by - ay) <= 2) || plus_hits_3x3(nbx - ax, nby - ay) ||
           plus_hits_3x3(nax - bx, nay - by) || (iabs_(nbx - nax) <= 2 && iabs_(nby - nay) <= 2);
  }
  static SM_HD bool box_hits_F(int ax, int ay, int nax, int nay, int bx, int by, int) {
    return plus_hits_box3(bx - ax, by - ay) || (iabs_(bx - nax) <= 4 && iabs_(by - nay) <= 4);
  }
};
template <> struct Foot<1> {
  static SM_HD bool in_range(int dx, int dy, int RA, int RB) { return iabs_(dx) <= RA + RB && iabs_(dy) <= RA + RB; }
  static SM_HD bool box_hits_M(int dx, int dy, int RB) {
    dx = iabs_(dx); dy = iabs_(dy);
    return (dx <= RB + 1 && dy <= RB) || (dx <= RB && dy <= RB + 1);
  }
  static SM_HD bool W_hits_M(int ibx, int iby, int nbx, int nby, int ax, int ay) {
    return plus_hits_5x5(ibx - ax, iby - ay) || plus_hits_5x5(nbx - ax, nby - ay);
  }
  static SM_HD bool c4(int dx, int dy) { return iabs_(dx) <= 4 && iabs_(dy) <= 4; }
  static SM_HD bool F_hits_F(int ax, int ay, int nax, int nay, int bx, int by, int nbx, int nby) {
    return c4(bx - ax, by - ay) || c4(nbx - ax, nby - ay) || c4(bx - nax, by - nay) || c4(nbx - nax, nby - nay);
  }
  static SM_HD bool box_hits_F(int ax, int ay, int nax, int nay, int bx, int by, int RB) {
    return (iabs_(bx - ax) <= RB + 2 && iabs_(by - ay) <= RB + 2) || (iabs_(bx - nax) <= RB + 2 && iabs_(by - nay) <= RB + 2);
  }
};
