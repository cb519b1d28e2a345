/* oracle/ASSUMPTIONS.h — every decision the CPU oracle makes that the reference's own tree does not pin.
 *
 * TEST INFRASTRUCTURE ONLY. Nothing under oracle/ is part of the product; only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs may load it.
 *
 * The reference (/root/reference, 894 lines of Rust) delegates the arithmetic of the hot path to third-party crates
 * that are NOT vendored and cannot be built here (no cargo/rustc, no network):
 *   sample 0.9.1 (pinned)            window::Windower, Hanning, Rectangle          Cargo.toml:8
 *   vox_box (git HEAD, UNPINNED)     spectrum::MFCC, hz_to_mel, mel_to_hz, dct     Cargo.toml:9
 *   rulinalg 0.4.2 (pinned)          utils::dot, Matrix::mean/variance             Cargo.toml:15
 *   rusty-machine 0.5.4 (pinned)     Standardizer, GaussianMixtureModel            Cargo.toml:14
 *   voting_experts (git HEAD, UNPINNED) cast_votes, split_string                   Cargo.toml:7
 * Each flag below fixes one recalled / published behaviour of those crates. Parity status per stage:
 *   pinned by a reference known answer : PCM decode scale (src/lib.rs:258), angular-distance clamp (src/sound.rs:612-615)
 *   fully specified by in-tree source  : cosine_sim/norm/at_distance (up to dot order A9), max_index, segment cutting,
 *                                        clone_from_dictionary, max_power/mean_mfccs (up to the windower rule A1)
 *   PARITY UNPINNED                    : MFCC values (A2-A4), GMM predict (A5-A6), Voting Experts (A7), DTW (A8; no
 *                                        reference counterpart exists at all)
 */
#ifndef SOUNDSYM_ORACLE_ASSUMPTIONS_H
#define SOUNDSYM_ORACLE_ASSUMPTIONS_H

/* A1  sample::window::Windower yields a window only while `bin <= remaining` and advances by `hop`
 *     => frames = n >= bin ? (n - bin)/hop + 1 : 0; the tail is dropped.  (call sites src/sound.rs:228-229, 245-246) */
#define ORC_A1_FULL_FRAMES_ONLY 1

/* A2  sample::window::Hanning: w(i) = 0.5*(1 - cos(2*pi*i/(bin-1)))  (symmetric, end points 0).
 *     Set to 0 for the periodic form (denominator bin). */
#define ORC_A2_HANN_SYMMETRIC 1

/* A3  vox_box MFCC: mel = 1125 ln(1+hz/700); ncoeffs+2 mel points spaced (hi-lo)/ncoeffs (top edge overshoots);
 *     bin_i = floor((N+1) hz_i / sr); un-normalised triangles (rise starts at weight 0, fall starts at 1);
 *     power spectrum |X|^2 when 1, magnitude |X| when 0; log10; DCT-II scaled by 2, k = 0..ncoeffs-1. */
#define ORC_A3_POWER_SPECTRUM 1

/* A4  energy floor before log10 so digital silence stays finite: log10(max(e, ORC_A4_ENERGY_FLOOR)).
 *     (tests/sample.wav has 1 859 all-zero frames; the reference as recalled would emit -inf there.) */
#define ORC_A4_ENERGY_FLOOR 1e-10

/* A5  rusty-machine Standardizer::transform = fit on the input, per-column mean and SAMPLE (n-1) variance,
 *     then (x - mean)/sqrt(var).  (src/lib.rs:57-58) */
#define ORC_A5_VARIANCE_DDOF 1

/* A6  GaussianMixtureModel::predict = membership weights
 *     w_ij = pi_j * exp(-0.5 (x_i-mu_j)^T Sigma_j^-1 (x_i-mu_j)) / sqrt(det Sigma_j), row-normalised, no (2 pi)^(d/2);
 *     symbol = max_index(row) with running max starting at (0, 0.0), strict '>' (src/sound.rs:486-495), so all-zero
 *     or NaN rows map to index 0 ('A'). */

/* A7  Voting Experts (Cohen & Adams): n-gram counts for lengths 1..depth+1; per length, frequency and boundary
 *     entropy are z-scored (population std; std == 0 -> z = 0); window of `depth` symbols slides one symbol at a
 *     time; inside each window the frequency expert votes for the split p in 1..depth-1 maximising
 *     z_f(w[..p]) + z_f(w[p..]) and the entropy expert for the split p in 1..depth maximising z_H(w[..p]);
 *     ties go to the EARLIEST p; votes has n+1 entries (votes[i] = boundary before symbol i).
 *     split_string cuts before symbol i (0 < i < n) iff votes[i] > votes[i-1] && votes[i] >= votes[i+1]
 *     && votes[i] >= threshold.  Chunks cover the whole string. */

/* A8  DTW (north-star extension; no reference function): local cost c(i,j) = sum_k (a_ik - b_jk)^2;
 *     D(0,0) = c(0,0); D(i,j) = c(i,j) + min(D(i-1,j), D(i,j-1), D(i-1,j-1)); no band;
 *     distance = D(Lq-1, Ld-1) / (Lq + Ld); empty segment -> +inf; winner = argmin (distance, index) lexicographic;
 *     NaN distances never win; top-k ascending. */

/* A9  rulinalg::utils::dot: 8 independent accumulators over chunks of 8, s = 0 + (p0+p4) + (p1+p5) + (p2+p6) +
 *     (p3+p7), then a scalar tail; products and sums are separate IEEE operations (no FMA contraction). */
#define ORC_A9_DOT_UNROLL8 1

#endif
