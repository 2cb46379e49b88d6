// Schedule of the host-buffer GEMM pipeline (host_pipeline.cuh): pure host arithmetic, no CUDA, so that it can be
// inspected (rtat_b200_host_schedule) and tested without a GPU.
//
// Column blocks of op(B) / C: at most 8, at least 256 columns each, multiples of 128; the last one is halved because
// its copy back to the host is the one transfer nothing can hide.  Every column block needs all of A, so a plain
// column-block pipeline idles the GPU while A crosses PCIe.  Phase 1 therefore runs the first `lead` column blocks
// K-panel by K-panel: panel p of A and the matching rows of op(B) under those blocks are copied, then multiplied
// (beta, then 1) while panel p + 1 is in flight.  `lead` is chosen so that those blocks' math takes about as long as
// copying A plus their own share of B.  Phase 2 is the column-block pipeline for the rest (A is resident by then):
// copy B_j, multiply, copy C_j back.  The first K-panel is split in two so that the math starts after half a panel.
#pragma once
#include <cstddef>

namespace rtat_b200 {

struct HostSchedule {
  static constexpr int MAX_BLOCKS = 8, MAX_PANELS = 8;
  static constexpr int BLOCK_CAP = MAX_BLOCKS + 6, PANEL_CAP = 2 * MAX_PANELS + 8;  // with the tapered tail / head
  int blocks = 0, lead = 0, npanels = 0, nb = 0;
  int col0[BLOCK_CAP] = {}, cols[BLOCK_CAP] = {};  // column blocks of op(B) / C
  int k0s[PANEL_CAP] = {}, kcs[PANEL_CAP] = {};    // K-panels of A
};

// elem: sizeof(T).  bytes_a: bytes of A as stored.  a_share: fraction of A that crosses this GPU's PCIe link, pcie_rate:
// bytes/s that link can expect, multi: the multi-GPU caller (takes a margin on `lead`, see below).  opt_panels /
// opt_lead: overrides for experiments (0 = the choice made here).
inline HostSchedule plan_host_schedule(int m, int n, int k, size_t elem, size_t bytes_a, double a_share, double pcie_rate, bool multi,
                                       int opt_panels, int opt_lead) {
  constexpr int MAX_BLOCKS = HostSchedule::MAX_BLOCKS, MAX_PANELS = HostSchedule::MAX_PANELS;
  HostSchedule sc;
  int nb = (n + MAX_BLOCKS - 1) / MAX_BLOCKS;
  nb = ((nb < 256 ? 256 : nb) + 127) / 128 * 128;
  sc.nb = nb;
  int blocks = (n + nb - 1) / nb;
  for (int j = 0; j < blocks; j++) {
    sc.col0[j] = j * nb;
    sc.cols[j] = (sc.col0[j] + nb > n) ? n - sc.col0[j] : nb;
  }
  // The copy back of the last block is the one transfer nothing can hide, so it is tapered — halved again and again down to
  // 256 columns (4096 -> 2048, 1024, 512, 256, 256) — each piece's copy back hides behind the next piece's product, which is
  // still several times longer (config 5: the exposed tail goes from 537 MB to 67 MB, 10 ms to 1.3 ms; where a product is
  // shorter than a copy the copies simply queue up, as they would have untapered).  Column blocks are rank-local, so the
  // multi-GPU caller tapers too (round 2's first version halved its last block once).
  while (blocks > 1 && sc.cols[blocks - 1] >= 512 && blocks < HostSchedule::BLOCK_CAP) {
    const int half = sc.cols[blocks - 1] / 2 / 128 * 128;
    sc.col0[blocks] = sc.col0[blocks - 1] + half;
    sc.cols[blocks] = sc.cols[blocks - 1] - half;
    sc.cols[blocks - 1] = half;
    blocks++;
  }
  int panels = 1, lead = 0;
  // The multi-GPU entry is collective: every rank must cut K into the SAME panels (the peers' "panel p has arrived" flags
  // are indexed by p), so for that caller the decision may depend only on what the collective contract fixes — m, k, lda,
  // transa (through bytes_a) and the world size — never on this rank's n_loc.  The single-GPU caller keeps the
  // "a single column block cannot be pipelined" shortcut.
  if ((multi || blocks > 1) && k >= 2048 && bytes_a >= (size_t(32) << 20)) {
    panels = k / 1024 < MAX_PANELS ? k / 1024 : MAX_PANELS;
    const double rate = elem == 8 ? 33e12 : 150e12;  // nominal: only the split point depends on them
    const double t_a = double(bytes_a) * a_share / pcie_rate, t_b = double(elem) * k * nb / pcie_rate, t_mm = 2.0 * m * nb * double(k) / rate;
    // in units of nb columns (leading blocks are full).  With several GPUs behind one host the link rate is the shaky
    // figure (they share the host's memory bandwidth) and a K-panel product is as efficient as a whole-K one at these
    // depths, so the multi-GPU caller gets a margin: one leading block at N = 4 measured 0.934 of the HBM-resident rate
    lead = t_mm > t_b ? (multi ? int(1.3 * t_a / (t_mm - t_b)) + 1 : int(t_a / (t_mm - t_b) + 0.5)) : blocks;
    if (lead < 1) lead = 1;
    // Copy-back bound problems (config 3: 98 MB of C leave in 1.8 ms, the whole product takes 1.45): the copies back must start
    // as early as possible, i.e. some column block must be COMPLETE early, and a leading block is complete only when all of A
    // is up.  Two leading blocks carry the K-panel phase, the rest follow whole-K with their copies back behind them (config 3
    // end to end 4.12 -> 3.52 ms with every block leading; 1 block: 3.58, 4 blocks: 3.63).
    if (!multi && double(elem) * m * n / pcie_rate > 0.75 * (2.0 * m * double(n) * k / rate) && lead > 2) lead = 2;
    if (opt_panels > 0) panels = opt_panels < MAX_PANELS ? opt_panels : MAX_PANELS;
    if (opt_lead > 0) lead = opt_lead;
    if (lead > blocks) lead = blocks;
  }
  const int kp = panels > 1 ? ((k + panels - 1) / panels + 15) / 16 * 16 : k;
  int npanels = 0;
  // Head of the single-GPU schedule: the first panel is what the GPU waits for, so it starts at 256 k and each panel may
  // exceed the first by `growth` times everything uploaded before it — growth = t_product / t_upload - 1 per k-row, the
  // largest step that never lets the product of panel p - 1 end before panel p has landed (config 5: 0.37, 1.5 ms instead
  // of 12 ms before the first product starts).  A balanced problem — `lead` above makes product and upload take equally
  // long, growth ~ 0.07 on config 2 — has no slack to grow into and its steady state stalls on every upload hiccup, so it
  // takes ONE MORE leading block when that buys a growth of 0.15 or more (config 2: 5 leading blocks, growth 0.24; end to
  // end 33.03 -> 32.3 ms, probe/perf_host2.py: the halved-first-panel schedule left the math waiting 0.85 ms for panel 2
  // and ran every panel product against an upload just as long).  Panels keep growing past the regular width kp (deeper
  // products, fewer beta = 1 passes over the leading blocks' C) up to 4 kp.
  double growth = 0.0;
  if (!multi && panels > 1 && lead > 0 && opt_panels <= 0) {
    const double rate = elem == 8 ? 33e12 : 150e12;
    auto growth_of = [&](int ld) {
      const double lead_cols = double(nb) * ld;
      const double t_up = (double(bytes_a) / k * a_share + double(elem) * lead_cols) / pcie_rate, t_mm = 2.0 * m * lead_cols / rate;
      return t_mm / t_up - 1.0;
    };
    growth = growth_of(lead);
    // the tapered tail keeps at least two full blocks' worth of columns for phase 2 (their products hide the copies back)
    if (growth < 0.15 && opt_lead == 0 && (lead + 3) * nb <= n && growth_of(lead + 1) >= 0.15) growth = growth_of(++lead);
    // experiments: host.panels = -g forces the growing head with a growth of at least g per cent
    if (opt_panels < 0 && growth < -0.01 * opt_panels) growth = -0.01 * opt_panels;
  }
  // Multi-GPU caller: the cuts are part of the collective contract, so the growth cannot follow this rank's column count.
  // `lead` above carries a 1.3x margin of product over upload, which a growth of 0.2 stays inside; the head then starts at
  // 256 k instead of half a regular panel (N = 8: 67 MB of this rank's share of A before the first product, 8 MB now).
  if (multi && panels > 1 && opt_panels == 0) growth = 0.2;
  if ((growth >= 0.15 || opt_panels < 0) && kp >= 1024) {
    if (growth > 1.0) growth = 1.0;
    const int wmax = 4 * kp;
    int done = 0;
    while (done < k && npanels < HostSchedule::PANEL_CAP - 1) {
      int w = npanels == 0 ? 256 : int(256 + growth * done) / 16 * 16;
      if (w > wmax) w = wmax;
      if (done + w > k || npanels == HostSchedule::PANEL_CAP - 2) w = k - done;
      // a sliver left over would be one more pass over C for nothing: the last panel takes it
      if (k - done - w < 256) w = k - done;
      sc.k0s[npanels] = done;
      sc.kcs[npanels++] = w;
      done += w;
    }
    sc.blocks = blocks;
    sc.lead = lead;
    sc.npanels = npanels;
    return sc;
  }
  for (int p0 = 0; p0 < k; p0 += kp) {
    sc.k0s[npanels] = p0;
    sc.kcs[npanels++] = p0 + kp > k ? k - p0 : kp;
  }
  if (panels > 1 && sc.kcs[0] >= 1024) {
    for (int p = npanels; p > 1; p--) { sc.k0s[p] = sc.k0s[p - 1]; sc.kcs[p] = sc.kcs[p - 1]; }
    const int half = sc.kcs[0] / 2 / 16 * 16;
    sc.k0s[1] = half;
    sc.kcs[1] = sc.kcs[0] - half;
    sc.kcs[0] = half;
    npanels++;
  }
  sc.blocks = blocks;
  sc.lead = lead;
  sc.npanels = npanels;
  return sc;
}

// Link model of the multi-GPU host entry (mg.cu): each of `world` ranks uploads 1 / world of A; the host side of one box
// sustains about 150 GB/s over all links, one link about 50 GB/s.
inline void mg_link_model(int world, double &a_share, double &pcie_rate) {
  a_share = 1.0 / world;
  pcie_rate = world > 3 ? 150e9 / world : 50e9;
}

}  // namespace rtat_b200
