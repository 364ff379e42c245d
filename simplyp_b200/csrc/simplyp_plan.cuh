// simplyp_b200 — placement of a cost-ordered, latency-bound ensemble on the SMs: the arithmetic of the plan
// (which virtual blocks form which list, where a member of cost rank r sits).  Pure integer functions shared by the
// launcher, cost_scatter_kernel and plan_claim (simplyp_kernels.cu) and by the host-side test harness.
//
// B virtual blocks of 32 members, 2 resident blocks per SM on n_sm SMs.  With Q = max(0, B - 2 n_sm) virtual blocks
// too many for the machine, nY = n_sm - Q, nP = min(nY, B - nY - 3Q), xb = nY + nP:
//   first-list t < nY        : heavy virtual block t;                    its second-list: virtual block nY + t (t < nP)
//   first-list nY + x, x < Q : light blocks xb+Q+x then B-1-x (chained); its second-list: light block xb+x
// The launch has one block per list; a block claims its list by where it lands: the first block to arrive on an SM
// takes the next first-list, the second one the second-list that belongs to it (plan_claim).  cost_scatter_kernel
// With the 3-blocks-per-SM build (`resident` = 3) nothing is chained: the launch has 3 n_sm blocks, all resident at
// once; the Q SMs whose first block drew a ticket >= nY hold the three light blocks of a slot side by side (third-list
// 2 n_sm + nY + x = block B-1-x), the third block to arrive on any other SM has no list and leaves.  A slot that runs
// two light blocks one after the other ends late once the light blocks are more than half as heavy as the heaviest
// (measured after the day tolerance became rate-dependent: 10.4 ms against 8.2-8.7 ms for the other SMs).
// cost_scatter_kernel
// lays the members out so that virtual blocks [0, nY) are the heaviest in descending order, [nY, nY+nP) their
// partners in ASCENDING order (the heaviest block shares its SM with the lightest partner) and [xb, B) the 3Q
// lightest blocks in descending order (the two blocks that share a slot are taken from the lightest 2Q, a heavier
// one with a lighter one; the block beside them from the Q above).  So the hardware never queues a block (a queued
// block starts only when the first resident one ends, 7.6 ms into a 14 ms run at 10^4 members), and the two light
// blocks that must share a slot run beside a third light block that ends about when the first of them does.
#pragma once
#include "simplyp_core.cuh"

namespace simplyp {

struct PlanShape {
  int n_sm, nY, nP, Q;
  int resident;                                            // blocks per SM the kernel build keeps resident: 2 or 3
  SP_HD int n_blocks() const { return nY + nP + 3 * Q; }
  SP_HD int n_lists() const { return resident == 3 ? n_blocks() : n_sm + nP + Q; }
  SP_HD int n_launch() const { return resident == 3 ? 3 * n_sm : n_lists(); }   // blocks of the launch
};

// Shape of the plan for B virtual blocks on n_sm SMs; false if the plan does not apply (B <= n_sm, or more than the
// lightest third of the resident set would have to be chained).
SP_HD bool plan_shape(long long B, int n_sm, PlanShape& p, int resident = 2) {
  if (n_sm <= 0 || B <= n_sm || B >= 3ll * n_sm) return false;
  const long long Q = B > 2ll * n_sm ? B - 2ll * n_sm : 0;
  const long long nY = n_sm - Q;
  if (nY <= 0) return false;
  long long nP = B - nY - 3 * Q;
  if (nP <= 0) return false;
  if (nP > nY) return false;                     // cannot happen for Q >= 0 (B <= 2 n_sm + Q), kept as a guard
  p.n_sm = n_sm; p.nY = (int)nY; p.nP = (int)nP; p.Q = (int)Q;
  p.resident = (resident == 3 && Q > 0) ? 3 : 2;
  return true;
}

// First virtual block of a list (-1: the list does not exist); lists [0, n_sm) are the first-lists, list n_sm + t is
// the second-list of first-list t, list 2 n_sm + t its third-list (resident = 3 only).
SP_HD int plan_list_head(const PlanShape& p, int list) {
  const int xb = p.nY + p.nP;
  if (list < 0) return -1;
  if (list < p.n_sm) return list < p.nY ? list : xb + p.Q + (list - p.nY);
  int t = list - p.n_sm;
  if (t < p.n_sm) {
    if (t < p.nY) return t < p.nP ? p.nY + t : -1;
    return xb + (t - p.nY);
  }
  t -= p.n_sm;
  if (p.resident != 3 || t < p.nY || t >= p.n_sm) return -1;
  return p.n_blocks() - 1 - (t - p.nY);
}
// virtual block that follows `vb` in its list, or -1
SP_HD int plan_list_next(const PlanShape& p, int vb) {
  if (p.resident == 3) return -1;
  const int x = vb - (p.nY + p.nP + p.Q);
  return (x >= 0 && x < p.Q) ? p.n_blocks() - 1 - x : -1;
}

// Member layout of a planned launch.  r = cost rank of a member (0 = heaviest); the result is its item index
// (virtual block = index / 32).
//  * the `solo` heaviest members each lead a warp of their own whose other seven quads hold light members (the
//    lightest of the partner region), so that the longest lock-step chain of the launch is one member's step count,
//    not the per-day maximum over eight similar heavy members;
//  * the blocks of the partner region [nY, nY + n_rev) are stored in reverse order (lightest first).
struct MemberLayout { int solo, nY, n_rev, fill_end; };

SP_HD MemberLayout member_layout(const PlanShape& p, int M, int solo) {
  MemberLayout L;
  if (8 * solo > 32 * p.nY || 7 * solo > 16 * p.nP || solo < 0) solo = 0;
  L.solo = solo;
  L.nY = p.nY;
  L.n_rev = p.nP - ((p.Q == 0 && M % 32 != 0) ? 1 : 0);      // a ragged last block stays last
  const long long pe = 32ll * (p.nY + p.nP);
  L.fill_end = (int)(pe < M ? pe : M);
  return L;
}

SP_HD int member_layout_index(int r, const MemberLayout& L) {
  const int K = L.solo, fill_begin = L.fill_end - 7 * K;
  if (r < K) return 8 * r;
  if (r >= fill_begin && r < L.fill_end) { const int t = r - fill_begin; return 8 * (t / 7) + 1 + t % 7; }
  if (r < fill_begin) r += 7 * K;
  const int blk = r >> 5;
  if (blk >= L.nY && blk < L.nY + L.n_rev) return ((L.nY + (L.n_rev - 1 - (blk - L.nY))) << 5) | (r & 31);
  return r;
}

}  // namespace simplyp
