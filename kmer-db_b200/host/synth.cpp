// Pattern-level synthetic database generator (ours; nothing like it exists in the reference).
//
// BASELINE.json's configs name "synthetic 5 Mbp bacterial genomes"; pushing 5-50 Gbp of FASTA
// through `build` on every benchmark run is impractical (SURVEY.md §7 "hard parts", §8d), so
// this simulates the reference's build *semantics* (PrefixKmerDb::addKmers,
// src/prefix_kmer_db.cpp:181-240,244-434) directly on k-mer RUNS instead of
// k-mers.  Model: samples form clusters; the first genome of a cluster is random, every later
// one is a copy of a uniformly chosen earlier member with i.i.d. substitutions at rate mu.  A
// substitution at base b replaces the k k-mers covering b by novel ones.  Random 18-mers do
// not collide at these sizes, so a k-mer's identity is (origin genome, origin position) and a
// maximal run of k-mers never cut by any mutation has a single membership set, i.e. a single
// pattern.  Runs are kept in a refinement tree (a cut leaf gets children); per sample we
// group the sample's leaves by pattern id and apply the extend-or-split rule of the
// reference (src/prefix_kmer_db.cpp:210-230).  Clusters share no k-mers, so they are
// simulated independently (in parallel) and merged in sample order.
// The output is a valid kmer-db trie: the unmodified reference's all2all accepts the .db
// written from it and the k-mer accounting is exact (tests/test_host.py).
#include <algorithm>
#include <cmath>
#include <cstring>
#include <thread>

#include "gamma.h"
#include "synth.h"

namespace kdbx {
namespace {

struct Rng {  // splitmix64
    uint64_t s;
    explicit Rng(uint64_t seed) : s(seed) {}
    uint64_t next() {
        uint64_t z = (s += 0x9E3779B97F4A7C15ull);
        z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
        z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
        return z ^ (z >> 31);
    }
    double uniform() { return ((next() >> 11) + 0.5) * (1.0 / 9007199254740992.0); }  // (0,1)
    uint64_t below(uint64_t n) { return (uint64_t)(uniform() * (double)n) % n; }
};

constexpr uint32_t kLeaf = 0xFFFFFFFFu;

struct ClusterSim {
    // run refinement tree
    std::vector<uint32_t> run_len, run_child, run_nchild;
    std::vector<int32_t> run_pat;
    // cluster-local patterns
    std::vector<int64_t> num_kmers;
    std::vector<int32_t> parent;
    std::vector<uint32_t> n, l, last, bits, born;
    std::vector<uint8_t> has_child;
    std::vector<uint64_t> pb_off;   // word offset of the pattern's gamma bits in arena
    std::vector<uint32_t> pb_cap;   // capacity in words
    std::vector<uint64_t> arena;
    // genomes as ordered run-id lists
    std::vector<std::vector<uint32_t>> genomes;
    uint64_t novel_kmers = 0;

    uint32_t new_run(uint32_t len, int32_t pat) {
        run_len.push_back(len); run_child.push_back(kLeaf); run_nchild.push_back(0);
        run_pat.push_back(pat);
        return (uint32_t)run_len.size() - 1;
    }
    int32_t new_pattern(int64_t kmers, int32_t par, uint32_t nn, uint32_t sid) {
        num_kmers.push_back(kmers); parent.push_back(par); n.push_back(nn); l.push_back(1);
        last.push_back(sid); bits.push_back(0); born.push_back(sid); has_child.push_back(0);
        pb_off.push_back(0); pb_cap.push_back(0);
        return (int32_t)num_kmers.size() - 1;
    }
    void append_sample(int32_t q, uint32_t sid) {  // pattern_t::expand (src/pattern.h:195-203)
        const uint32_t delta = sid - last[q];
        const uint32_t need_bits = bits[q] + gamma_code_len(delta);
        const uint32_t need_words = (need_bits + 63) / 64 + 1;
        if (need_words > pb_cap[q]) {
            const uint32_t ncap = std::max<uint32_t>(2, std::max(need_words, pb_cap[q] * 2));
            const uint64_t noff = arena.size();
            arena.resize(noff + ncap, 0);
            if (pb_cap[q]) std::memcpy(&arena[noff], &arena[pb_off[q]], (size_t)pb_cap[q] * 8);
            pb_off[q] = noff; pb_cap[q] = ncap;
        }
        gamma_put(&arena[pb_off[q]], bits[q], delta);
        last[q] = sid; ++n[q]; ++l[q];
    }

    unsigned team = 1;  // threads working on this cluster

    // f(tid, begin, end) over [0,n) in `team` contiguous slices (slice t on thread t)
    template <class F>
    void par_for(size_t n, F&& f) const {
        if (team <= 1 || n < 8192) { f(0u, (size_t)0, n); return; }
        const size_t per = (n + team - 1) / team;
        std::vector<std::thread> th;
        th.reserve(team);
        for (unsigned t = 0; t < team; ++t)
            th.emplace_back([&f, t, per, n]() { const size_t b = std::min(n, t * per); f(t, b, std::min(n, b + per)); });
        for (auto& x : th) x.join();
    }

    // Replace every run of a genome by its current leaves (in genomic order).
    void refine(std::vector<uint32_t>& g) {
        std::vector<std::vector<uint32_t>> part(team);
        par_for(g.size(), [&](unsigned tid, size_t b, size_t e) {
            std::vector<uint32_t>& out = part[tid];
            std::vector<uint32_t> stack;
            out.reserve((e - b) + (e - b) / 8);
            for (size_t i = b; i < e; ++i) {
                const uint32_t r = g[i];
                if (run_child[r] == kLeaf) { out.push_back(r); continue; }
                stack.clear(); stack.push_back(r);
                while (!stack.empty()) {
                    const uint32_t x = stack.back(); stack.pop_back();
                    if (run_child[x] == kLeaf) { out.push_back(x); continue; }
                    for (uint32_t c = run_nchild[x]; c-- > 0;) stack.push_back(run_child[x] + c);
                }
            }
        });
        size_t total = 0;
        for (auto& v : part) total += v.size();
        g.clear(); g.reserve(total);
        for (auto& v : part) g.insert(g.end(), v.begin(), v.end());
    }

    struct CutOut {  // what one slice of the cut phase produces
        std::vector<uint32_t> list;      // run ids; new runs as kNewFlag | local index
        std::vector<uint32_t> new_len;
        std::vector<int32_t> new_pat;
        std::vector<uint32_t> split_run, split_first, split_cnt;  // cut leaves and their children (local indices)
        uint64_t len_sum = 0;
    };
    static constexpr uint32_t kNewFlag = 0x80000000u;

    void run(const SynthParams& sp, const std::vector<uint32_t>& sids, uint64_t seed) {
        Rng rng(seed);
        const uint64_t L = sp.genome_kmers;
        const uint32_t k = sp.k;
        genomes.resize(sids.size());
        std::vector<uint32_t> touched;
        std::vector<uint32_t> cnt;       // k-mers of the current sample per pattern
        std::vector<int32_t> remap;
        std::vector<std::pair<uint64_t, uint64_t>> dead;  // destroyed k-mer intervals [s,e)
        const double log1m = std::log1p(-sp.mutation_rate);
        const int32_t root_marker = -2;  // leaves that are novel in this sample

        for (size_t m = 0; m < sids.size(); ++m) {
            const uint32_t sid = sids[m];
            std::vector<uint32_t>& gi = genomes[m];
            uint64_t novel = 0;
            if (m == 0) {
                novel = L;
            } else {
                const size_t j = (size_t)rng.below(m);
                refine(genomes[j]);
                const std::vector<uint32_t>& gj = genomes[j];
                // destroyed k-mer intervals from substitution positions (base coordinates)
                dead.clear();
                const uint64_t bases = L + k - 1;
                uint64_t b = 0;
                for (;;) {
                    const double u = rng.uniform();
                    const uint64_t gap = (uint64_t)std::floor(std::log(u) / log1m);
                    b += gap;
                    if (b >= bases) break;
                    const uint64_t s = b >= (uint64_t)(k - 1) ? b - (k - 1) : 0;
                    const uint64_t e = std::min<uint64_t>(b + 1, L);
                    if (!dead.empty() && s <= dead.back().second) dead.back().second = std::max(dead.back().second, e);
                    else if (s < e) dead.emplace_back(s, e);
                    ++b;
                }
                // walk the leaves of genome j, cutting those a destroyed interval overlaps
                std::vector<CutOut> outs(team);
                par_for(gj.size(), [&](unsigned tid, size_t b0, size_t e0) {  // slice lengths
                    uint64_t sum = 0;
                    for (size_t i = b0; i < e0; ++i) sum += run_len[gj[i]];
                    outs[tid].len_sum = sum;
                });
                std::vector<uint64_t> start(team + 1, 0);
                for (unsigned t = 0; t < team; ++t) start[t + 1] = start[t] + outs[t].len_sum;
                par_for(gj.size(), [&](unsigned tid, size_t b0, size_t e0) {
                    CutOut& o = outs[tid];
                    o.list.reserve((e0 - b0) + (e0 - b0) / 4);
                    uint64_t pos = start[tid];
                    size_t di = (size_t)(std::lower_bound(dead.begin(), dead.end(), pos,
                                                          [](const std::pair<uint64_t, uint64_t>& iv, uint64_t p) { return iv.second <= p; }) -
                                         dead.begin());
                    auto add_run = [&](uint32_t len, int32_t pat) {
                        o.new_len.push_back(len); o.new_pat.push_back(pat);
                        return (uint32_t)o.new_len.size() - 1;
                    };
                    for (size_t idx = b0; idx < e0; ++idx) {
                        const uint32_t r = gj[idx];
                        const uint64_t len = run_len[r], end = pos + len;
                        while (di < dead.size() && dead[di].second <= pos) ++di;
                        if (di == dead.size() || dead[di].first >= end) {  // untouched leaf
                            o.list.push_back(r); pos = end; continue;
                        }
                        // pieces: alternate kept / destroyed inside [pos,end); children contiguous
                        const uint32_t first_child = (uint32_t)o.new_len.size();
                        uint32_t nchild = 0;
                        uint64_t cur = pos;
                        size_t d = di;
                        const int32_t pat = run_pat[r];
                        while (cur < end) {
                            if (d < dead.size() && dead[d].first < end) {
                                const uint64_t ds = std::max(dead[d].first, cur), de = std::min(dead[d].second, end);
                                if (ds > cur) { o.list.push_back(kNewFlag | add_run((uint32_t)(ds - cur), pat)); ++nchild; }  // kept
                                add_run((uint32_t)(de - ds), pat); ++nchild;  // destroyed here, lives on elsewhere
                                cur = de;
                                if (dead[d].second <= end) ++d; else break;
                            } else {
                                o.list.push_back(kNewFlag | add_run((uint32_t)(end - cur), pat)); ++nchild;
                                cur = end;
                            }
                        }
                        o.split_run.push_back(r); o.split_first.push_back(first_child); o.split_cnt.push_back(nchild);
                        pos = end;
                    }
                });
                // append the new runs slice by slice (== serial order), then patch ids
                std::vector<uint32_t> base(team + 1, (uint32_t)run_len.size());
                size_t list_total = 0;
                std::vector<size_t> list_at(team + 1, 0);
                for (unsigned t = 0; t < team; ++t) {
                    base[t + 1] = base[t] + (uint32_t)outs[t].new_len.size();
                    list_at[t + 1] = list_at[t] + outs[t].list.size();
                }
                list_total = list_at[team];
                const size_t nruns = base[team];
                run_len.resize(nruns); run_pat.resize(nruns); run_child.resize(nruns, kLeaf); run_nchild.resize(nruns, 0);
                gi.resize(list_total);
                par_for(team, [&](unsigned, size_t tb, size_t te) {
                    for (size_t t = tb; t < te; ++t) {
                        const CutOut& o = outs[t];
                        std::copy(o.new_len.begin(), o.new_len.end(), run_len.begin() + base[t]);
                        std::copy(o.new_pat.begin(), o.new_pat.end(), run_pat.begin() + base[t]);
                        for (size_t i = 0; i < o.list.size(); ++i) {
                            const uint32_t v = o.list[i];
                            gi[list_at[t] + i] = (v & kNewFlag) ? base[t] + (v & ~kNewFlag) : v;
                        }
                        for (size_t i = 0; i < o.split_run.size(); ++i) {
                            run_child[o.split_run[i]] = base[t] + o.split_first[i];
                            run_nchild[o.split_run[i]] = o.split_cnt[i];
                        }
                    }
                });
                for (auto& iv : dead) novel += iv.second - iv.first;
            }
            // Novel k-mers (the whole first genome; afterwards one per destroyed k-mer, since
            // substitutions keep the length).  They all share one pattern — this sample's root —
            // and their position inside the genome does not matter to the build rule: append.
            for (uint64_t left = novel; left;) {
                const uint32_t len = (uint32_t)std::min<uint64_t>(left, 0x7FFFFFFFu);  // run lengths are 32-bit
                gi.push_back(new_run(len, root_marker));
                left -= len;
            }

            // ---- addKmers(sample sid): group by pattern, extend or split ----
            if (cnt.size() < num_kmers.size() + 1) cnt.resize(num_kmers.size() + (num_kmers.size() >> 2) + 1024, 0);
            {
                std::vector<std::vector<uint32_t>> tl(team);
                par_for(gi.size(), [&](unsigned tid, size_t b0, size_t e0) {
                    for (size_t i = b0; i < e0; ++i) {
                        const uint32_t r = gi[i];
                        const int32_t q = run_pat[r];
                        if (q < 0) continue;
                        if (__atomic_fetch_add(&cnt[q], run_len[r], __ATOMIC_RELAXED) == 0) tl[tid].push_back((uint32_t)q);
                    }
                });
                touched.clear();
                for (auto& v : tl) touched.insert(touched.end(), v.begin(), v.end());
                std::sort(touched.begin(), touched.end());  // pattern numbering independent of the team size
            }
            const size_t first_new = num_kmers.size();
            int32_t root = -1;
            if (novel) root = new_pattern((int64_t)novel, -1, 1, sid);
            if (remap.size() < first_new) remap.resize(first_new + (first_new >> 2) + 1024, -1);
            for (uint32_t q : touched) {
                const uint32_t c = cnt[q];
                if ((int64_t)c == num_kmers[q] && !has_child[q]) {
                    append_sample((int32_t)q, sid);
                    remap[q] = (int32_t)q;
                } else {
                    const int32_t r = new_pattern((int64_t)c, (int32_t)q, n[q] + 1, sid);
                    num_kmers[q] -= c; has_child[q] = 1;
                    remap[q] = r;
                }
            }
            par_for(gi.size(), [&](unsigned, size_t b0, size_t e0) {
                for (size_t i = b0; i < e0; ++i) {
                    const uint32_t r = gi[i];
                    const int32_t q = run_pat[r];
                    run_pat[r] = (q < 0) ? root : remap[q];
                }
            });
            for (uint32_t q : touched) { cnt[q] = 0; remap[q] = -1; }
            novel_kmers += novel;
        }
        // free what the merge does not need
        genomes.clear(); genomes.shrink_to_fit();
        run_len.clear(); run_len.shrink_to_fit(); run_child.clear(); run_child.shrink_to_fit();
        run_nchild.clear(); run_nchild.shrink_to_fit(); run_pat.clear(); run_pat.shrink_to_fit();
    }
};

}  // namespace

void synth_generate(const SynthParams& sp_in, Trie& t) {
    SynthParams sp = sp_in;
    if (sp.num_samples == 0 || sp.num_clusters == 0 || sp.genome_kmers == 0 || sp.k == 0 ||
        !(sp.mutation_rate > 0.0 && sp.mutation_rate < 1.0))
        throw std::runtime_error("synth_generate: bad parameters");
    sp.num_clusters = std::min(sp.num_clusters, sp.num_samples);
    const uint32_t N = sp.num_samples, C = sp.num_clusters;

    // sample -> cluster assignment
    std::vector<uint32_t> cluster_of(N);
    std::vector<std::vector<uint32_t>> members(C);
    std::vector<uint32_t> first_of(C + 1, N);   // contiguous order: cluster c holds the samples [first_of[c], first_of[c+1])
    {
        // unequal clusters: weight 1 + skew * u_c with u_c in [-1, 1) fixed by the cluster index; at least one sample each
        std::vector<double> cum(C + 1, 0.0);
        for (uint32_t c = 0; c < C; ++c) {
            const double u = (double)((c * 2654435761u) % 1000u) / 500.0 - 1.0;
            cum[c + 1] = cum[c] + 1.0 + (sp.cluster_skew > 0.0 && sp.cluster_skew < 1.0 ? sp.cluster_skew * u : 0.0);
        }
        first_of[0] = 0;
        for (uint32_t c = 1; c < C; ++c) {
            uint32_t b = sp.cluster_skew > 0.0 && sp.cluster_skew < 1.0 ? (uint32_t)(cum[c] / cum[C] * N + 0.5) : (uint32_t)(((uint64_t)c * N + C - 1) / C);
            b = std::max(b, first_of[c - 1] + 1);
            first_of[c] = std::min(b, N - (C - c));
        }
    }
    for (uint32_t s = 0, cc = 0; s < N; ++s) {
        while (!sp.interleaved && s >= first_of[cc + 1]) ++cc;
        const uint32_t c = sp.interleaved ? (s % C) : cc;
        cluster_of[s] = c; members[c].push_back(s);
    }
    std::vector<ClusterSim> sims(C);
    {
        // clusters run side by side; spare cores form a team inside each cluster
        const unsigned total = sp.threads > 0 ? (unsigned)sp.threads : std::max(1u, std::thread::hardware_concurrency());
        const unsigned nt = std::min<unsigned>(total, C);
        const unsigned team = std::max(1u, std::min(16u, total / nt));
        std::vector<std::thread> th;
        for (unsigned w = 0; w < nt; ++w)
            th.emplace_back([&, w]() {
                for (uint32_t c = w; c < C; c += nt) {
                    sims[c].team = team;
                    sims[c].run(sp, members[c], sp.seed * 0x100000001B3ull + c + 1);
                }
            });
        for (auto& x : th) x.join();
    }

    // ---- merge clusters in sample order: patterns born at sample s keep their relative order
    uint64_t P = 1;
    for (auto& s : sims) P += s.num_kmers.size();
    t.hdr = DbHeader();
    t.hdr.kmer_length = sp.k;
    t.hdr.num_hashtables = (uint64_t)1 << std::max<int>(8, (int)sp.k * 2 - 32);  // src/prefix_kmer_db.cpp:54-62
    t.sample_names.resize(N); t.sample_kmers.assign(N, sp.genome_kmers);
    for (uint32_t s = 0; s < N; ++s) {
        char buf[32]; std::snprintf(buf, sizeof buf, "g%06u", s);
        t.sample_names[s] = buf;
    }
    t.num_kmers.resize(P); t.parent_id.resize(P); t.n.resize(P); t.l.resize(P);
    t.last.resize(P); t.bits.resize(P); t.payload_off.resize(P);
    t.num_kmers[0] = 0; t.parent_id[0] = -1; t.n[0] = 0; t.l[0] = 0; t.last[0] = 0; t.bits[0] = 0;
    t.payload_off[0] = 0;

    std::vector<std::vector<uint64_t>> gid(C);  // cluster-local pattern -> global id
    for (uint32_t c = 0; c < C; ++c) gid[c].resize(sims[c].num_kmers.size());
    std::vector<size_t> cursor(C, 0);
    uint64_t next = 1, payload_words = 0;
    for (uint32_t s = 0; s < N; ++s) {
        const uint32_t c = cluster_of[s];
        ClusterSim& sim = sims[c];
        size_t& i = cursor[c];
        while (i < sim.num_kmers.size() && sim.born[i] == s) { gid[c][i] = next++; ++i; }
    }
    uint64_t kmers_total = 0;
    for (uint32_t c = 0; c < C; ++c) {
        ClusterSim& sim = sims[c];
        kmers_total += sim.novel_kmers;
        for (size_t i = 0; i < sim.num_kmers.size(); ++i) {
            const uint64_t g = gid[c][i];
            t.num_kmers[g] = sim.num_kmers[i];
            t.parent_id[g] = sim.parent[i] < 0 ? -1 : (int64_t)gid[c][sim.parent[i]];
            t.n[g] = sim.n[i]; t.l[g] = sim.l[i]; t.last[g] = sim.last[i]; t.bits[g] = sim.bits[i];
            payload_words += Trie::payload_words_for_bits(sim.bits[i]);
        }
    }
    t.hdr.kmers_count = kmers_total;
    t.payload.resize(payload_words, 0);
    // payload in global pattern order
    {
        std::vector<std::pair<uint32_t, uint32_t>> where(P, {0u, 0u});
        for (uint32_t c = 0; c < C; ++c)
            for (size_t i = 0; i < sims[c].num_kmers.size(); ++i) where[gid[c][i]] = {c, (uint32_t)i};
        uint64_t off = 0;
        for (uint64_t g = 1; g < P; ++g) {
            const ClusterSim& sim = sims[where[g].first];
            const uint32_t i = where[g].second;
            t.payload_off[g] = off;
            const uint64_t words = Trie::payload_words_for_bits(sim.bits[i]);
            if (words) {
                const uint64_t have = std::min<uint64_t>(words, sim.pb_cap[i]);
                std::memcpy(t.payload.data() + off, &sim.arena[sim.pb_off[i]], have * 8);
                off += words;
            }
        }
    }
    t.build_compact();
}

}  // namespace kdbx
