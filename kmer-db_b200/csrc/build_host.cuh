// Host side of the device builder (included by kdbx.cu after `struct kdbx_builder`): capacity
// management, the per-sample launch sequence and the final export.  See build.cuh for the kernels.
#pragma once

#define BCK(call)                                                                              \
    do {                                                                                       \
        cudaError_t e__ = (call);                                                              \
        if (e__ != cudaSuccess)                                                                \
            return b->ctx->fail(e__ == cudaErrorMemoryAllocation ? KDBX_ERR_NOMEM : KDBX_ERR_CUDA, \
                                "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e__), __FILE__, __LINE__); \
    } while (0)

inline int bits_for(uint64_t v) { int b = 1; while (b < 64 && (v >> b)) ++b; return b; }
inline uint64_t pow2_at_least(uint64_t v) { uint64_t c = 1; while (c < v) c <<= 1; return c; }

// grows a buffer to `need` bytes keeping its first `used` bytes
int grow_keep(kdbx_builder* b, DevBuf& buf, size_t used, size_t need) {
    if (need <= buf.bytes) return KDBX_OK;
    const size_t want = std::max(need + need / 8 + 256, buf.bytes * 2);
    void* np = nullptr;
    BCK(cudaMalloc(&np, want));
    if (used && buf.p) BCK(cudaMemcpyAsync(np, buf.p, used, cudaMemcpyDeviceToDevice, b->ctx->stream));
    BCK(cudaStreamSynchronize(b->ctx->stream));
    if (buf.p) cudaFree(buf.p);
    buf.p = np; buf.bytes = want;
    return KDBX_OK;
}

int builder_reserve_patterns(kdbx_builder* b, uint64_t need) {
    if (need <= b->pat_cap) return KDBX_OK;
    const uint64_t cap = std::max<uint64_t>(need + need / 4 + 1024, b->pat_cap * 2);
    const uint64_t P = b->P;
    if (int rc = grow_keep(b, b->num_kmers, P * 8, cap * 8)) return rc;
    if (int rc = grow_keep(b, b->parent, P * 8, cap * 8)) return rc;
    if (int rc = grow_keep(b, b->n, P * 4, cap * 4)) return rc;
    if (int rc = grow_keep(b, b->l, P * 4, cap * 4)) return rc;
    if (int rc = grow_keep(b, b->last, P * 4, cap * 4)) return rc;
    if (int rc = grow_keep(b, b->born, P * 4, cap * 4)) return rc;
    if (int rc = grow_keep(b, b->is_parent, P * 4, cap * 4)) return rc;
    b->pat_cap = cap;
    return KDBX_OK;
}

int builder_reserve_events(kdbx_builder* b, uint64_t need) {
    if (need <= b->ev_cap) return KDBX_OK;
    const uint64_t cap = std::max<uint64_t>(need + need / 4 + 1024, b->ev_cap * 2);
    if (int rc = grow_keep(b, b->ev_pat, b->ev_count * 4, cap * 4)) return rc;
    if (int rc = grow_keep(b, b->ev_sample, b->ev_count * 4, cap * 4)) return rc;
    b->ev_cap = cap;
    return KDBX_OK;
}

// the table keeps its load at or below one half: probes stay short and inserts always terminate
int builder_reserve_table(kdbx_builder* b, uint64_t incoming) {
    const uint64_t need = b->filled + incoming;
    if (b->cap && need * 2 <= b->cap) return KDBX_OK;
    cudaStream_t st = b->ctx->stream;
    const uint64_t cap = std::max<uint64_t>(pow2_at_least(need * 5 / 2 + 16), (uint64_t)1 << 16);
    void *nk = nullptr, *nv = nullptr;
    BCK(cudaMalloc(&nk, cap * 8));
    if (cudaMalloc(&nv, cap * 4) != cudaSuccess) { cudaFree(nk); cudaGetLastError(); return b->ctx->fail(KDBX_ERR_NOMEM, "k-mer table of %llu slots does not fit in device memory", (unsigned long long)cap); }
    BCK(cudaMemsetAsync(nk, 0xFF, cap * 8, st));
    BCK(cudaMemsetAsync(nv, 0, cap * 4, st));
    if (b->cap && b->filled)
        k_rehash<<<blocks_for(b->cap, 256), 256, 0, st>>>(b->keys.as<unsigned long long>(), b->vals.as<uint32_t>(), b->cap,
                                                           static_cast<unsigned long long*>(nk), static_cast<uint32_t*>(nv), cap - 1);
    BCK(cudaStreamSynchronize(st));
    BCK(cudaGetLastError());
    b->keys.release(); b->vals.release();
    b->keys.p = nk; b->keys.bytes = cap * 8;
    b->vals.p = nv; b->vals.bytes = cap * 4;
    b->cap = cap;
    b->table_growths += 1;
    return KDBX_OK;
}

int builder_fail_flag(kdbx_builder* b, int flag) {
    switch (flag) {
        case 0: return KDBX_OK;
        case 1: return b->ctx->fail(KDBX_ERR_ARG, "kdbx_builder_add_kmers: k-mers must ascend strictly and fit the k-mer length");
        case 6: return b->ctx->fail(KDBX_ERR_ARG, "adopted k-mer table points at a pattern that does not exist");
        case 7: return b->ctx->fail(KDBX_ERR_STATE, "builder: the event log does not match the patterns (internal error)");
        case 8: return b->ctx->fail(KDBX_ERR_STATE, "builder: a k-mer prefix lies outside the database's table range");
        default: return b->ctx->fail(KDBX_ERR_STATE, "builder: device error flag %d", flag);
    }
}

// The sample's unique k-mers are in b->uniq[0..count): tables, grouping, extend-or-split.
int builder_add_unique(kdbx_builder* b, uint64_t count) {
    kdbx_ctx* ctx = b->ctx;
    cudaStream_t st = ctx->stream;
    if (b->finished) return ctx->fail(KDBX_ERR_STATE, "builder already finished");
    if (b->num_samples == 0xFFFFFFFEu) return ctx->fail(KDBX_ERR_ARG, "too many samples");
    const uint32_t sample = b->num_samples;
    if (count == 0) { b->num_samples += 1; return KDBX_OK; }   // registered, nothing to add (src/prefix_kmer_db.cpp:256-260)
    if (count >= 0xFFFFFFFFull) return ctx->fail(KDBX_ERR_ARG, "a sample may hold fewer than 2^32 k-mers");
    if (b->P + count >= 0x7FFFFFFFull) return ctx->fail(KDBX_ERR_ARG, "too many patterns");
    if (int rc = builder_reserve_table(b, count)) return rc;
    if (int rc = builder_reserve_patterns(b, b->P + count)) return rc;
    if (int rc = builder_reserve_events(b, b->ev_count + count)) return rc;
    BCK(b->slot_of.ensure(count * 8)); BCK(b->pid.ensure(count * 4)); BCK(b->pid2.ensure(count * 4));
    BCK(b->idx.ensure(count * 4)); BCK(b->idx2.ensure(count * 4)); BCK(b->head.ensure(count * 4));
    BCK(b->head_incl.ensure(count * 4)); BCK(b->run_start.ensure((count + 1) * 4)); BCK(b->split.ensure(count * 4));
    BCK(b->split_incl.ensure(count * 4)); BCK(b->run_tag.ensure(count * 4));
    BuildPerSample* ps = b->ps.as<BuildPerSample>();
    BCK(cudaMemsetAsync(&ps->inserted, 0, offsetof(BuildPerSample, unique_kmers), st));  // keeps unique_kmers and err
    const unsigned grid = blocks_for(count, 256);
    k_find_or_insert<<<grid, 256, 0, st>>>(b->uniq.as<unsigned long long>(), count, b->keys.as<unsigned long long>(), b->vals.as<uint32_t>(),
                                           b->cap - 1, b->slot_of.as<unsigned long long>(), b->pid.as<uint32_t>(), b->idx.as<uint32_t>(), ps);
    {
        size_t tmp = 0;
        const int end_bit = std::min(32, bits_for(b->P));
        BCK(cub::DeviceRadixSort::SortPairs(nullptr, tmp, b->pid.as<uint32_t>(), b->pid2.as<uint32_t>(), b->idx.as<uint32_t>(), b->idx2.as<uint32_t>(),
                                            count, 0, end_bit, st));
        BCK(ctx->cub_tmp.ensure(tmp));
        BCK(cub::DeviceRadixSort::SortPairs(ctx->cub_tmp.p, tmp, b->pid.as<uint32_t>(), b->pid2.as<uint32_t>(), b->idx.as<uint32_t>(),
                                            b->idx2.as<uint32_t>(), count, 0, end_bit, st));
    }
    k_run_heads<<<grid, 256, 0, st>>>(b->pid2.as<uint32_t>(), count, b->head.as<uint32_t>());
    {
        size_t tmp = 0;
        BCK(cub::DeviceScan::InclusiveSum(nullptr, tmp, b->head.as<uint32_t>(), b->head_incl.as<uint32_t>(), count, st));
        BCK(ctx->cub_tmp.ensure(tmp));
        BCK(cub::DeviceScan::InclusiveSum(ctx->cub_tmp.p, tmp, b->head.as<uint32_t>(), b->head_incl.as<uint32_t>(), count, st));
    }
    k_run_starts<<<grid, 256, 0, st>>>(b->head.as<uint32_t>(), b->head_incl.as<uint32_t>(), count, b->run_start.as<uint32_t>(), ps);
    k_decide<<<grid, 256, 0, st>>>(b->pid2.as<uint32_t>(), b->run_start.as<uint32_t>(), count, ps, b->num_kmers.as<long long>(),
                                   b->is_parent.as<uint32_t>(), b->split.as<uint32_t>());
    {
        size_t tmp = 0;
        BCK(cub::DeviceScan::InclusiveSum(nullptr, tmp, b->split.as<uint32_t>(), b->split_incl.as<uint32_t>(), count, st));
        BCK(ctx->cub_tmp.ensure(tmp));
        BCK(cub::DeviceScan::InclusiveSum(ctx->cub_tmp.p, tmp, b->split.as<uint32_t>(), b->split_incl.as<uint32_t>(), count, st));
    }
    k_apply_runs<<<grid, 256, 0, st>>>(b->pid2.as<uint32_t>(), b->run_start.as<uint32_t>(), count, ps, b->split.as<uint32_t>(),
                                       b->split_incl.as<uint32_t>(), sample, b->P, b->ev_count, b->num_kmers.as<long long>(),
                                       b->parent.as<long long>(), b->n.as<uint32_t>(), b->l.as<uint32_t>(), b->last.as<uint32_t>(),
                                       b->born.as<uint32_t>(), b->is_parent.as<uint32_t>(), b->ev_pat.as<uint32_t>(),
                                       b->ev_sample.as<uint32_t>(), b->run_tag.as<uint32_t>());
    k_repoint<<<grid, 256, 0, st>>>(b->head_incl.as<uint32_t>(), b->idx2.as<uint32_t>(), count, b->run_tag.as<uint32_t>(),
                                    b->slot_of.as<unsigned long long>(), b->vals.as<uint32_t>());
    b->launches += 11;
    BCK(cudaMemcpyAsync(b->h_ps, ps, sizeof(BuildPerSample), cudaMemcpyDeviceToHost, st));
    BCK(cudaStreamSynchronize(st));
    BCK(cudaGetLastError());
    if (b->h_ps->err) { const int f = b->h_ps->err; cudaMemsetAsync(&ps->err, 0, 4, st); return builder_fail_flag(b, f); }
    if (b->h_ps->num_runs == 0 || b->h_ps->num_splits > b->h_ps->num_runs || b->h_ps->inserted > count)
        return ctx->fail(KDBX_ERR_STATE, "builder: inconsistent per-sample counters (internal error)");
    b->filled += b->h_ps->inserted;
    b->P += b->h_ps->num_splits;
    b->ev_count += b->h_ps->num_runs - b->h_ps->num_splits;
    b->num_samples += 1;
    b->total_kmers += count;
    return KDBX_OK;
}

__global__ void k_strip_sentinel(const unsigned long long* __restrict__ uniq, const unsigned long long* __restrict__ nsel,
                                 unsigned long long sentinel, BuildPerSample* __restrict__ ps) {
    unsigned long long c = *nsel;
    if (c && uniq[c - 1] >= sentinel) --c;
    ps->unique_kmers = c;
}

int builder_add_sequence(kdbx_builder* b, const char* symbols, uint64_t len, uint64_t* unique_kmers) {
    kdbx_ctx* ctx = b->ctx;
    cudaStream_t st = ctx->stream;
    if (b->finished) return ctx->fail(KDBX_ERR_STATE, "builder already finished");
    if (len && !symbols) return ctx->fail(KDBX_ERR_ARG, "kdbx_builder_add_sequence: symbols is NULL");
    BCK(cudaSetDevice(ctx->device));
    uint64_t count = 0;
    if (len >= b->alpha.k) {
        BCK(b->seq.ensure(len + 16)); BCK(b->raw.ensure(len * 8)); BCK(b->sorted.ensure(len * 8)); BCK(b->uniq.ensure(len * 8));
        BCK(cudaMemcpyAsync(b->seq.p, symbols, len, cudaMemcpyHostToDevice, st));
        k_extract_kmers<<<blocks_for(len, 256), 256, 0, st>>>(b->seq.as<uint8_t>(), len, b->d_alpha.as<BuildAlphabet>(), b->raw.as<unsigned long long>());
        size_t tmp = 0;
        const int end_bit = (int)b->sentinel_bit + 1;
        BCK(cub::DeviceRadixSort::SortKeys(nullptr, tmp, b->raw.as<unsigned long long>(), b->sorted.as<unsigned long long>(), len, 0, end_bit, st));
        BCK(ctx->cub_tmp.ensure(tmp));
        BCK(cub::DeviceRadixSort::SortKeys(ctx->cub_tmp.p, tmp, b->raw.as<unsigned long long>(), b->sorted.as<unsigned long long>(), len, 0, end_bit, st));
        unsigned long long* nsel = b->nsel.as<unsigned long long>();
        BCK(cub::DeviceSelect::Unique(nullptr, tmp, b->sorted.as<unsigned long long>(), b->uniq.as<unsigned long long>(), nsel, len, st));
        BCK(ctx->cub_tmp.ensure(tmp));
        BCK(cub::DeviceSelect::Unique(ctx->cub_tmp.p, tmp, b->sorted.as<unsigned long long>(), b->uniq.as<unsigned long long>(), nsel, len, st));
        BuildPerSample* ps = b->ps.as<BuildPerSample>();
        k_strip_sentinel<<<1, 1, 0, st>>>(b->uniq.as<unsigned long long>(), nsel, b->alpha.sentinel, ps);
        b->launches += 5;
        BCK(cudaMemcpyAsync(b->h_ps, ps, sizeof(BuildPerSample), cudaMemcpyDeviceToHost, st));
        BCK(cudaStreamSynchronize(st));
        BCK(cudaGetLastError());
        count = b->h_ps->unique_kmers;
    }
    if (unique_kmers) *unique_kmers = count;
    return builder_add_unique(b, count);
}

int builder_add_kmers(kdbx_builder* b, const uint64_t* kmers, uint64_t count) {
    kdbx_ctx* ctx = b->ctx;
    cudaStream_t st = ctx->stream;
    if (b->finished) return ctx->fail(KDBX_ERR_STATE, "builder already finished");
    if (count && !kmers) return ctx->fail(KDBX_ERR_ARG, "kdbx_builder_add_kmers: kmers is NULL");
    BCK(cudaSetDevice(ctx->device));
    if (count) {
        BCK(b->uniq.ensure(count * 8));
        BCK(cudaMemcpyAsync(b->uniq.p, kmers, count * 8, cudaMemcpyHostToDevice, st));
        BuildPerSample* ps = b->ps.as<BuildPerSample>();
        k_check_sorted<<<blocks_for(count, 256), 256, 0, st>>>(b->uniq.as<unsigned long long>(), count, b->alpha.sentinel, &ps->err);
        b->launches += 1;
        // refuse bad input BEFORE it touches the table
        BCK(cudaMemcpyAsync(b->h_ps, ps, sizeof(BuildPerSample), cudaMemcpyDeviceToHost, st));
        BCK(cudaStreamSynchronize(st));
        if (b->h_ps->err) { const int f = b->h_ps->err; BCK(cudaMemsetAsync(&ps->err, 0, 4, st)); return builder_fail_flag(b, f); }
    }
    return builder_add_unique(b, count);
}

// Continue the database staged on the context (kdbx_load_patterns + kdbx_load_hashtables): `build -extend`.
int builder_adopt(kdbx_builder* b) {
    kdbx_ctx* ctx = b->ctx;
    cudaStream_t st = ctx->stream;
    if (b->num_samples || b->P != 1 || b->filled) return ctx->fail(KDBX_ERR_STATE, "kdbx_builder_adopt: the builder already holds samples");
    if (!ctx->loaded || !ctx->tables_loaded) return ctx->fail(KDBX_ERR_STATE, "kdbx_builder_adopt: stage the database first (kdbx_load_patterns + kdbx_load_hashtables)");
    if (ctx->num_tables != b->num_tables) return ctx->fail(KDBX_ERR_ARG, "kdbx_builder_adopt: the database has %llu k-mer tables, these parameters imply %llu",
                                                           (unsigned long long)ctx->num_tables, (unsigned long long)b->num_tables);
    if (int rcw = require_full_window(ctx, "kdbx_builder_adopt")) return rcw;
    BCK(cudaSetDevice(ctx->device));
    Plan pl;
    if (int rc = make_plan(ctx, pl)) return rc;
    uint32_t launches = 0;
    const int rc = prepare(ctx, pl, launches);   // nodes + decoded local ids
    if (rc < 0) return rc;
    BCK(cudaStreamSynchronize(st));
    if (int rc2 = check_device_error(ctx)) return rc2;
    const uint64_t P = ctx->P;
    if (int r = builder_reserve_patterns(b, P + 1024)) return r;
    BCK(cudaMemcpyAsync(b->num_kmers.p, ctx->num_kmers.p, P * 8, cudaMemcpyDeviceToDevice, st));
    BCK(cudaMemcpyAsync(b->parent.p, ctx->parent.p, P * 8, cudaMemcpyDeviceToDevice, st));
    BCK(cudaMemcpyAsync(b->n.p, ctx->n.p, P * 4, cudaMemcpyDeviceToDevice, st));
    BCK(cudaMemcpyAsync(b->l.p, ctx->l.p, P * 4, cudaMemcpyDeviceToDevice, st));
    BCK(cudaMemcpyAsync(b->last.p, ctx->last.p, P * 4, cudaMemcpyDeviceToDevice, st));
    BCK(cudaMemsetAsync(b->is_parent.p, 0, P * 4, st));
    BCK(b->eoff.ensure((P + 1) * 8));
    {
        cub::CountingInputIterator<uint64_t> idx(0);
        cub::TransformInputIterator<uint64_t, EventsOf, cub::CountingInputIterator<uint64_t>> it(idx, EventsOf{ctx->l.as<uint32_t>(), P});
        if (int r = scan_exclusive(ctx, it, b->eoff.as<uint64_t>(), P + 1)) return r;
    }
    uint64_t events = 0;
    BCK(cudaMemcpyAsync(&events, b->eoff.as<uint64_t>() + P, 8, cudaMemcpyDeviceToHost, st));
    BCK(cudaStreamSynchronize(st));
    if (int r = builder_reserve_events(b, events + 1024)) return r;
    k_adopt_patterns<<<blocks_for(P, 256), 256, 0, st>>>(P, ctx->nodes.as<Node>(), ctx->loc.as<uint32_t>(), b->eoff.as<uint64_t>(),
                                                          b->born.as<uint32_t>(), b->is_parent.as<uint32_t>(), b->ev_pat.as<uint32_t>(),
                                                          b->ev_sample.as<uint32_t>());
    b->P = P;
    b->ev_count = events;
    // k-mer tables: every used raw slot {suffix, pattern} of bucket t is the k-mer (t << 32 | suffix)
    uint64_t total_slots = 0;
    BCK(cudaMemcpyAsync(&total_slots, ctx->slot_off.as<uint64_t>() + ctx->num_tables, 8, cudaMemcpyDeviceToHost, st));
    BCK(cudaStreamSynchronize(st));
    if (int r = builder_reserve_table(b, total_slots)) return r;   // upper bound of the k-mers coming in
    BuildPerSample* ps = b->ps.as<BuildPerSample>();
    BCK(cudaMemsetAsync(ps, 0, sizeof(BuildPerSample), st));
    k_adopt_tables<<<blocks_for(total_slots, 256), 256, 0, st>>>(total_slots, ctx->num_tables, ctx->slot_off.as<uint64_t>(),
                                                                  ctx->slots.as<unsigned long long>(), b->keys.as<unsigned long long>(),
                                                                  b->vals.as<uint32_t>(), b->cap - 1, P, ps);
    b->launches += launches + 3;
    BCK(cudaMemcpyAsync(b->h_ps, ps, sizeof(BuildPerSample), cudaMemcpyDeviceToHost, st));
    BCK(cudaStreamSynchronize(st));
    BCK(cudaGetLastError());
    if (b->h_ps->err) return builder_fail_flag(b, b->h_ps->err);
    b->filled = b->h_ps->inserted;
    b->num_samples = ctx->N;
    return KDBX_OK;
}

int builder_finish(kdbx_builder* b, kdbx_build_result* out) {
    kdbx_ctx* ctx = b->ctx;
    cudaStream_t st = ctx->stream;
    if (!out) return ctx->fail(KDBX_ERR_ARG, "kdbx_builder_finish: result is NULL");
    BCK(cudaSetDevice(ctx->device));
    const uint64_t P = b->P, E = b->ev_count, T = b->num_tables;
    if (!b->finished) {
        ctx->ev_used = 0;
        cudaEvent_t ev0 = ctx->event();
        BuildPerSample* ps = b->ps.as<BuildPerSample>();
        BCK(cudaMemsetAsync(ps, 0, sizeof(BuildPerSample), st));
        // events by pattern; the radix sort is stable, so a pattern's samples stay in ascending order
        BCK(b->ev_pat2.ensure((E + 1) * 4)); BCK(b->ev_sample2.ensure((E + 1) * 4));
        if (E) {
            size_t tmp = 0;
            const int end_bit = std::min(32, bits_for(P));
            BCK(cub::DeviceRadixSort::SortPairs(nullptr, tmp, b->ev_pat.as<uint32_t>(), b->ev_pat2.as<uint32_t>(), b->ev_sample.as<uint32_t>(),
                                                b->ev_sample2.as<uint32_t>(), E, 0, end_bit, st));
            BCK(ctx->cub_tmp.ensure(tmp));
            BCK(cub::DeviceRadixSort::SortPairs(ctx->cub_tmp.p, tmp, b->ev_pat.as<uint32_t>(), b->ev_pat2.as<uint32_t>(), b->ev_sample.as<uint32_t>(),
                                                b->ev_sample2.as<uint32_t>(), E, 0, end_bit, st));
        }
        BCK(b->eoff.ensure((P + 1) * 8)); BCK(b->poff.ensure((P + 1) * 8)); BCK(b->bits.ensure(P * 4));
        cub::CountingInputIterator<uint64_t> idx(0);
        {
            cub::TransformInputIterator<uint64_t, EventsOf, cub::CountingInputIterator<uint64_t>> it(idx, EventsOf{b->l.as<uint32_t>(), P});
            if (int r = scan_exclusive(ctx, it, b->eoff.as<uint64_t>(), P + 1)) return r;
        }
        uint64_t events = 0;
        BCK(cudaMemcpyAsync(&events, b->eoff.as<uint64_t>() + P, 8, cudaMemcpyDeviceToHost, st));
        BCK(cudaStreamSynchronize(st));
        if (events != E) return ctx->fail(KDBX_ERR_STATE, "builder: %llu logged extensions but the patterns hold %llu (internal error)",
                                          (unsigned long long)E, (unsigned long long)events);
        k_gamma_bits<<<blocks_for(P, 128), 128, 0, st>>>(P, b->l.as<uint32_t>(), b->born.as<uint32_t>(), b->last.as<uint32_t>(), b->eoff.as<uint64_t>(),
                                                          b->ev_pat2.as<uint32_t>(), b->ev_sample2.as<uint32_t>(), b->bits.as<uint32_t>(), &ps->err);
        {
            cub::TransformInputIterator<uint64_t, PayloadWords, cub::CountingInputIterator<uint64_t>> it(idx, PayloadWords{b->bits.as<uint32_t>(), P});
            if (int r = scan_exclusive(ctx, it, b->poff.as<uint64_t>(), P + 1)) return r;
        }
        uint64_t words = 0;
        BCK(cudaMemcpyAsync(&words, b->poff.as<uint64_t>() + P, 8, cudaMemcpyDeviceToHost, st));
        BCK(cudaStreamSynchronize(st));
        BCK(b->payload.ensure((words + 2) * 8));
        BCK(cudaMemsetAsync(b->payload.p, 0, (words + 2) * 8, st));
        k_gamma_encode<<<blocks_for(P, 128), 128, 0, st>>>(P, b->l.as<uint32_t>(), b->born.as<uint32_t>(), b->eoff.as<uint64_t>(),
                                                            b->ev_sample2.as<uint32_t>(), b->poff.as<uint64_t>(), b->payload.as<unsigned long long>());
        b->payload_words = words;
        // the reference's raw prefix tables
        BCK(b->tfilled.ensure((T + 1) * 8)); BCK(b->slot_off.ensure((T + 1) * 8));
        BCK(cudaMemsetAsync(b->tfilled.p, 0, (T + 1) * 8, st));
        if (b->cap)
            k_prefix_count<<<blocks_for(b->cap, 256), 256, 0, st>>>(b->keys.as<unsigned long long>(), b->cap, T, b->tfilled.as<unsigned long long>(), &ps->err);
        {
            cub::TransformInputIterator<uint64_t, TableCapacity, cub::CountingInputIterator<uint64_t>> it(idx, TableCapacity{b->tfilled.as<unsigned long long>(), T});
            if (int r = scan_exclusive(ctx, it, b->slot_off.as<uint64_t>(), T + 1)) return r;
        }
        uint64_t total_slots = 0;
        BCK(cudaMemcpyAsync(&total_slots, b->slot_off.as<uint64_t>() + T, 8, cudaMemcpyDeviceToHost, st));
        BCK(cudaMemcpyAsync(b->h_ps, ps, sizeof(BuildPerSample), cudaMemcpyDeviceToHost, st));
        BCK(cudaStreamSynchronize(st));
        if (b->h_ps->err) return builder_fail_flag(b, b->h_ps->err);
        BCK(b->slots.ensure(total_slots * 8 + 8));
        k_fill_u64<<<blocks_for(total_slots, 256), 256, 0, st>>>(b->slots.as<unsigned long long>(), total_slots, 0x7FFFFFFFull << 32);
        if (b->cap)
            k_export_tables<<<blocks_for(b->cap, 256), 256, 0, st>>>(b->keys.as<unsigned long long>(), b->vals.as<uint32_t>(), b->cap,
                                                                      b->slot_off.as<uint64_t>(), b->slots.as<unsigned long long>());
        b->total_slots = total_slots;
        b->launches += 10;
        cudaEvent_t ev1 = ctx->event();
        BCK(cudaStreamSynchronize(st));
        BCK(cudaGetLastError());
        b->ms_finish = elapsed(ev0, ev1);
        b->finished = true;
    }
    std::memset(out, 0, sizeof *out);
    out->num_patterns = P; out->payload_words = b->payload_words; out->num_tables = T; out->total_slots = b->total_slots;
    out->kmers_count = b->filled; out->num_samples = b->num_samples; out->kernel_launches = b->launches;
    out->table_capacity = b->cap; out->table_growths = b->table_growths; out->sum_local_samples = E + (P ? P - 1 : 0);
    out->ms_finish = b->ms_finish;
    return KDBX_OK;
}

int builder_export(kdbx_builder* b, const kdbx_build_arrays* a) {
    kdbx_ctx* ctx = b->ctx;
    cudaStream_t st = ctx->stream;
    if (!b->finished) return ctx->fail(KDBX_ERR_STATE, "kdbx_builder_export: call kdbx_builder_finish first");
    if (!a) return ctx->fail(KDBX_ERR_ARG, "kdbx_builder_export: arrays is NULL");
    BCK(cudaSetDevice(ctx->device));
    const uint64_t P = b->P, T = b->num_tables;
    auto copy = [&](void* dst, const DevBuf& src, size_t bytes) -> cudaError_t {
        if (!dst || !bytes) return cudaSuccess;
        return cudaMemcpyAsync(dst, src.p, bytes, cudaMemcpyDeviceToHost, st);
    };
    BCK(copy(a->num_kmers, b->num_kmers, P * 8)); BCK(copy(a->parent_id, b->parent, P * 8));
    BCK(copy(a->num_samples_full, b->n, P * 4)); BCK(copy(a->num_local_samples, b->l, P * 4));
    BCK(copy(a->last_sample_id, b->last, P * 4)); BCK(copy(a->num_bits, b->bits, P * 4));
    BCK(copy(a->payload_off, b->poff, P * 8)); BCK(copy(a->payload, b->payload, b->payload_words * 8));
    BCK(copy(a->slot_off, b->slot_off, (T + 1) * 8)); BCK(copy(a->slots, b->slots, b->total_slots * 8));
    BCK(copy(a->table_filled, b->tfilled, T * 8));
    BCK(cudaStreamSynchronize(st));
    return KDBX_OK;
}
