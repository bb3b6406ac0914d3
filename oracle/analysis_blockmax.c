/*
 * ORACLE-SIDE ANALYSIS — TEST INFRASTRUCTURE ONLY, never part of the product path.
 *
 * How many of maxscore's block decodes on the NON-ESSENTIAL lists would block-level score bounds save?  ds2i's wand
 * data holds one bound per list (max_term_weight, wand_data.hpp:35-50); a bound per 128-posting block could be derived
 * at index-load time (like the block directory) and checked before a block is decoded.  This tool runs the plain-C
 * restatement of maxscore_query (ds2i_oracle.c, queries.hpp:478-591) over a query log twice per query — as the reference
 * does, and with the extra check — and counts the docs blocks decoded by next_geq on non-essential lists.  Results of both
 * runs must agree (the check is rank-safe).  Used for DESIGN.md §8, not by any test.
 *
 *   analysis_blockmax <block index type> <index> <wand> <queries> [max_queries]
 */
#define DS2I_ORACLE_NO_MAIN
#include "ds2i_oracle.c"

static float** g_blockmax;       /* per term: max doc_term_weight of every block (lazily computed) */
static uint32_t** g_blockfmax;   /* per term: max freq of every block; entry [blocks] = max freq of the list */

static const float* blockmax_of(const block_index* ix, const wand_data* w, uint64_t term) {
    if (g_blockmax[term]) return g_blockmax[term];
    static enumerator e;
    enum_open(&e, ix, term);
    float* bm = (float*)calloc(e.blocks, sizeof(float));
    uint32_t* fm = (uint32_t*)calloc(e.blocks + 1, sizeof(uint32_t));
    for (uint32_t i = 0; i < e.n; ++i) {
        uint32_t f = (uint32_t)enum_freq(&e);
        float s = doc_term_weight(f, rdf(w->norm_lens + 4 * (uint64_t)e.cur_docid));
        if (s > bm[i / 128]) bm[i / 128] = s;
        if (f > fm[i / 128]) fm[i / 128] = f;
        if (f > fm[e.blocks]) fm[e.blocks] = f;
        enum_next(&e);
    }
    g_blockfmax[term] = fm;
    return g_blockmax[term] = bm;
}

typedef struct { uint64_t probe_decodes, essential_decodes, skipped, scored; } counters;

/* maxscore_query with counters; use_blockmax adds the per-block check in front of every non-essential next_geq */
static uint64_t maxscore_counted(const block_index* ix, const wand_data* w, const uint32_t* terms, size_t n, topk_queue* topk,
                                 int use_blockmax, counters* c) {
    topk->size = 0;
    if (!n) return 0;
    size_t m = setup(ix, w, terms, n), ord[MAXT];
    for (size_t i = 0; i < m; ++i) ord[i] = i;
    sort_idx(ord, m, 1);
    float ub[MAXT];
    ub[0] = g_enums[ord[0]].max_weight;
    for (size_t i = 1; i < m; ++i) ub[i] = ub[i - 1] + g_enums[ord[i]].max_weight;
    const float* bmax[MAXT];
    const uint32_t* fmax[MAXT];
    for (size_t i = 0; i < m; ++i) { bmax[i] = use_blockmax ? blockmax_of(ix, w, g_enums[ord[i]].term) : 0; fmax[i] = use_blockmax ? g_blockfmax[g_enums[ord[i]].term] : 0; }
    if (use_blockmax) setup(ix, w, terms, n);            /* blockmax_of used a scratch enumerator only, but keep cursors fresh */
    c->essential_decodes += m;                            /* block 0 of every list */
    uint64_t non_essential = 0, cur_doc = ix->num_docs;
    for (size_t i = 0; i < m; ++i) if (g_enums[i].e.cur_docid < cur_doc) cur_doc = g_enums[i].e.cur_docid;
    while (non_essential < m && cur_doc < ix->num_docs) {
        float score = 0, norm_len = rdf(w->norm_lens + 4 * cur_doc);
        uint64_t next_doc = ix->num_docs;
        for (size_t i = non_essential; i < m; ++i) {
            scored_enum* en = &g_enums[ord[i]];
            if (en->e.cur_docid == cur_doc) {
                score += en->q_weight * doc_term_weight(enum_freq(&en->e), norm_len);
                uint32_t b = en->e.cur_block;
                enum_next(&en->e);
                if (en->e.cur_block != b) c->essential_decodes += 1;
            }
            if (en->e.cur_docid < next_doc) next_doc = en->e.cur_docid;
        }
        c->scored += 1;
        int dead = 0;
        for (size_t i = non_essential - 1; i + 1 > 0; --i) {
            if (!topk_would_enter(topk, score + ub[i])) { dead = 1; break; }
            scored_enum* en = &g_enums[ord[i]];
            enumerator* e = &en->e;
            if (use_blockmax >= 3) {
                /* per-document bound from the LIST's largest freq and the document's own length: sum over lists i..0 */
                float rest = 0;
                for (size_t j = 0; j <= i; ++j) rest += g_enums[ord[j]].q_weight * doc_term_weight(fmax[j][g_enums[ord[j]].e.blocks], norm_len);
                if (!topk_would_enter(topk, score + rest)) { c->skipped += 1; dead = 1; break; }
            }
            if ((use_blockmax == 1 || use_blockmax == 2 || use_blockmax == 4) && cur_doc > e->cur_block_max && cur_doc <= block_max(e, e->blocks - 1)) {
                uint32_t b = e->cur_block + 1;
                while (block_max(e, b) < cur_doc) ++b;    /* the block next_geq would decode (block_posting_list.hpp:129-137) */
                float bound = use_blockmax == 1 ? en->q_weight * bmax[i][b] : en->q_weight * doc_term_weight(fmax[i][b], norm_len);
                float below = i ? ub[i - 1] : 0.f;
                if (use_blockmax == 4) { below = 0; for (size_t j = 0; j < i; ++j) below += g_enums[ord[j]].q_weight * doc_term_weight(fmax[j][g_enums[ord[j]].e.blocks], norm_len); }
                if (!topk_would_enter(topk, score + below + bound)) { c->skipped += 1; dead = 1; break; }
            }
            uint32_t before = e->cur_block;
            enum_next_geq(e, cur_doc);
            if (e->cur_block != before) c->probe_decodes += 1;
            if (e->cur_docid == cur_doc) score += en->q_weight * doc_term_weight(enum_freq(e), norm_len);
        }
        if (!dead || !use_blockmax) {
            if (topk_insert(topk, score))
                while (non_essential < m && !topk_would_enter(topk, ub[non_essential])) non_essential += 1;
        }
        cur_doc = next_doc;
    }
    return topk->size;
}

int main(int argc, char** argv) {
    if (argc < 5) { fprintf(stderr, "usage: analysis_blockmax <type> <index> <wand> <queries> [max_queries]\n"); return 1; }
    uint8_t *ibytes, *wbytes; size_t in, wn;
    if (load_file(argv[2], &ibytes, &in) || load_file(argv[3], &wbytes, &wn)) { perror("open"); return 1; }
    block_index ix;
    if (open_index(&ix, argv[1], ibytes)) { fprintf(stderr, "block index types only\n"); return 1; }
    wand_data w;
    open_wand(&w, wbytes);
    query_log q = read_queries(argv[4]);
    size_t nq = argc > 5 ? (size_t)strtoull(argv[5], 0, 10) : q.n;
    if (nq > q.n) nq = q.n;
    g_blockmax = (float**)calloc(ix.size, sizeof(float*));
    g_blockfmax = (uint32_t**)calloc(ix.size, sizeof(uint32_t*));
    static const char* names[5] = {"reference", "block max weight", "block max freq x document length", "list max freq x document length",
                                   "block max freq + list max freqs below, x document length"};
    counters c[5];
    memset(c, 0, sizeof c);
    topk_queue ta, tb;
    ta.k = tb.k = 10;
    uint64_t mismatches = 0;
    for (size_t i = 0; i < nq; ++i) {
        const uint32_t* t = q.terms + q.begin[i];
        size_t n = q.begin[i + 1] - q.begin[i];
        uint64_t ra = maxscore_counted(&ix, &w, t, n, &ta, 0, &c[0]);
        for (int mode = 1; mode < 5; ++mode) {
            uint64_t rb = maxscore_counted(&ix, &w, t, n, &tb, mode, &c[mode]);
            if (ra != rb || memcmp(ta.q, tb.q, 4 * ta.size)) ++mismatches;
        }
    }
    printf("{\"queries\": %zu, \"result_mismatches\": %llu", nq, (unsigned long long)mismatches);
    for (int mode = 0; mode < 5; ++mode)
        printf(", \"%s\": {\"probe_block_decodes\": %llu, \"essential_block_decodes\": %llu, \"probes_refused\": %llu}", names[mode],
               (unsigned long long)c[mode].probe_decodes, (unsigned long long)c[mode].essential_decodes, (unsigned long long)c[mode].skipped);
    printf("}\n");
    return 0;
}
