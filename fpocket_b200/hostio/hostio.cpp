// libfpocket_hostio: native, re-entrant host I/O either side of libfpocket_cuda (include/fpocket_hostio.h).
//
//   reader  : reference src/rpdb.c:785-1478 (rpdb_open + rpdb_read, both atom lists) in ONE pass over the file's bytes, followed by the
//             flattening the C shim does (fpocket_b200/shim/search_pocket_cuda.c: shim_flatten / shim_job_prepare)
//   writers : reference src/fpout.c:55-213, src/writepocket.c:262-640, src/write_visu.c:56-290, src/writepdb.c:78-216,
//             src/voronoi.c:1437-1557, fed from fpk_result instead of c_lst_pockets
//   pipeline: the -F loop of src/fpmain.c:61-76 as reader pool -> detect callback (fpk_detect_batch) -> writer pool
//
// Everything the reference's parser does with its fixed 83-byte line buffer is kept, including what it reads from a PREVIOUS line when a
// record is short (the buffer is never cleared, rpdb.c:1064), because that decides which bytes land in the output files.
#include "fpocket_hostio.h"

#include <atomic>
#include <cerrno>
#include <chrono>
#include <cmath>
#include <condition_variable>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <deque>
#include <functional>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include <fcntl.h>
#include <sys/stat.h>
#include <sys/types.h>
#include <unistd.h>

namespace {

#include "tables.inc"
#include "viz_templates.inc"

constexpr int LINE_MAX_CHARS = 81;      // fgets(buf, M_PDB_LINE_LEN + 2, f) hands out at most 81 characters (rpdb.h:32, rpdb.c:832)

struct AtomRec {                        // the fields of s_atm (headers/atom.h:29-77) the path and the writers use
    char type[8], name[8], res_name[8], chain[4], symbol[4];
    int32_t id, res_id, charge;
    char insert, aloc;
    float x, y, z, occ, bf, radius, en;
};

double now_s() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

// ---- the C library calls the reference makes on slices of its line buffer ------------------------------------------------------------
// atoi(p) with a NUL planted at p[len] (rpdb.c:283-287): leading blanks, sign, digits; stops at the slice end or an earlier NUL
int slice_atoi(const char *p, int len)
{
    int i = 0;
    while (i < len && p[i] && (p[i] == ' ' || (p[i] >= '\t' && p[i] <= '\r'))) i++;
    bool neg = false;
    if (i < len && (p[i] == '-' || p[i] == '+')) { neg = p[i] == '-'; i++; }
    long long v = 0;
    while (i < len && p[i] >= '0' && p[i] <= '9') { v = v * 10 + (p[i] - '0'); if (v > 0x7fffffffLL + 1) { v = 0x7fffffffLL + 1; } i++; }
    if (neg) v = -v;
    if (v > 0x7fffffffLL) v = 0x7fffffffLL;
    if (v < -0x80000000LL) v = -0x80000000LL;
    return (int)v;
}
// (float)atof(p) with a NUL planted at p[len] (rpdb.c:632-660). Plain decimals with <= 15 digits take the exact route
// (integer mantissa / power of ten is the correctly rounded double); anything else goes to strtod on a copy of the slice.
float slice_atof(const char *p, int len)
{
    int i = 0;
    while (i < len && p[i] && (p[i] == ' ' || (p[i] >= '\t' && p[i] <= '\r'))) i++;
    int j = i;
    bool neg = false;
    if (j < len && (p[j] == '-' || p[j] == '+')) { neg = p[j] == '-'; j++; }
    unsigned long long m = 0;
    int nd = 0, frac = 0;
    bool dot = false, simple = true;
    int k = j;
    for (; k < len && p[k]; k++) {
        const char c = p[k];
        if (c >= '0' && c <= '9') { m = m * 10 + (unsigned)(c - '0'); nd++; if (dot) frac++; }
        else if (c == '.' && !dot) dot = true;
        else break;
    }
    if (nd == 0 || nd > 15) simple = false;
    if (simple && k < len && p[k]) {
        const char c = p[k];            // what follows the digits must not continue a number strtod would accept (exponent, hex, inf/nan)
        if (c == 'e' || c == 'E' || c == 'x' || c == 'X' || c == 'p' || c == 'P') simple = false;
    }
    if (simple) {
        static const double P10[16] = {1e0, 1e1, 1e2, 1e3, 1e4, 1e5, 1e6, 1e7, 1e8, 1e9, 1e10, 1e11, 1e12, 1e13, 1e14, 1e15};
        double d = (double)m / P10[frac];
        return (float)(neg ? -d : d);
    }
    char tmp[24];
    int n = 0;
    while (n < len && n < 23 && p[n]) { tmp[n] = p[n]; n++; }
    tmp[n] = 0;
    return (float)atof(tmp);
}
// strncpy(dst, p, len); dst[len] = 0 (dst zero-filled beyond the copied bytes, as strncpy does)
void slice_strncpy(char *dst, const char *p, int len)
{
    int i = 0;
    for (; i < len && p[i]; i++) dst[i] = p[i];
    for (; i <= len; i++) dst[i] = 0;
}
void str_trim(char *s)                  // utils.c:518-531: blanks only
{
    int len = (int)strlen(s);
    while (len > 0 && s[len - 1] == ' ') s[--len] = 0;
    int lead = 0;
    while (lead < len && s[lead] == ' ') lead++;
    if (lead) { memmove(s, s + lead, (size_t)(len - lead) + 1); memset(s + (len - lead) + 1, 0, (size_t)lead); }   // the bytes behind the terminator end up NUL there too
}

// ---- element tables (pertable.c) ------------------------------------------------------------------------------------------------------
const ElemRow *elem_lookup(const char *symbol)         // pte_get_vdw_ray / pte_get_enegativity: first row matching (upper, lower)
{
    const char a0 = (char)toupper((unsigned char)symbol[0]), a1 = (char)tolower((unsigned char)symbol[1]);
    if (a1 == 0) {                                     // the one-letter symbols, in table order (first match wins there too)
        switch (a0) {
        case 'C': return &ELEMENTS[6];
        case 'N': return &ELEMENTS[7];
        case 'O': return &ELEMENTS[8];
        case 'S': return &ELEMENTS[16];
        case 'H': return &ELEMENTS[1];
        case 'P': return &ELEMENTS[15];
        }
    }
    for (const ElemRow &e : ELEMENTS) if (e.c0 == a0 && e.c1 == a1) return &e;
    return nullptr;
}
bool in_std_res(const char *r)          // pertable.c element_in_std_res: strncmp(res_name, X, 3)
{
    static const char *N[26] = {"GLY", "LEU", "ILE", "TRP", "MET", "SER", "THR", "LYS", "ARG", "ASN", "GLN", "GLU", "ASP", "CYS", "PRO", "HIS",
                                "TYR", "PHE", "VAL", "ALA", "HIE", "HID", "HIP", "HSD", "HSE", "HSP"};
    for (const char *n : N) if (!strncmp(r, n, 3)) return true;
    return false;
}
bool in_nucl_acid(const char *r)
{
    static const char *N[9] = {"dG", "dC", "dT", "dA", "A", "C", "T", "G", "U"};
    for (const char *n : N) if (!strncmp(r, n, 3)) return true;
    return false;
}
bool first_letter_in(const char *s, const char *letters)      // is_valid_prot_element / is_valid_nucl_acid_element with ignore_case
{
    if (!s[0]) return false;
    char t[8];
    snprintf(t, sizeof t, "%s", s);
    str_trim(t);
    const char c = (char)tolower((unsigned char)t[0]);
    for (const char *q = letters; *q; q++) if (c == (char)tolower((unsigned char)*q)) return true;
    return false;
}
bool is_element_name(const char *s)                           // is_valid_element(str, 1): the whole trimmed string is a symbol
{
    if (!s[0]) return false;
    char t[8];
    snprintf(t, sizeof t, "%s", s);
    str_trim(t);
    t[0] = (char)tolower((unsigned char)t[0]);
    if (t[0]) t[1] = (char)tolower((unsigned char)t[1]);
    for (const ElemRow &e : ELEMENTS) {
        const char u[3] = {(char)tolower((unsigned char)e.c0), (char)tolower((unsigned char)e.c1), 0};
        if (!strcmp(t, u)) return true;
    }
    return false;
}
void guess_element(const char *aname, char *element, const char *res_name)      // rpdb.c:410-461
{
    char tmp[8];
    snprintf(tmp, sizeof tmp, "%s", aname);
    str_trim(tmp);
    const char *pt = tmp;
    if (isdigit((unsigned char)tmp[0])) pt++;
    if (in_std_res(res_name)) {
        if (first_letter_in(pt, "HCNOSP")) { element[0] = pt[0]; element[1] = 0; return; }
    } else if (in_nucl_acid(res_name)) {
        if (first_letter_in(pt, "HCNOP")) { element[0] = pt[0]; element[1] = 0; return; }
    } else if (is_element_name(pt)) {
        element[0] = pt[0]; element[1] = pt[0] ? pt[1] : 0; element[2] = 0;       // strcpy of a one- or two-letter symbol
        return;
    }
    element[0] = pt[0];
    element[1] = pt[0] ? pt[1] : 0;
    element[2] = 0;
}
bool keep_hetatm(const char *resb)      // rpdb.c:952-961: first three bytes equal to an entry of ST_keep_hetatm
{
    if (!resb[0] || !resb[1]) return false;     // the one-letter entries never match: the byte the reference compares lies beyond their terminator
    const uint32_t key = ((uint32_t)(unsigned char)resb[0] << 16) | ((uint32_t)(unsigned char)resb[1] << 8) | (uint32_t)(unsigned char)resb[2];
    int lo = 0, hi = (int)(sizeof KEEP_HETATM / sizeof KEEP_HETATM[0]) - 1;
    while (lo <= hi) {
        const int mid = (lo + hi) >> 1;
        if (KEEP_HETATM[mid] == key) return true;
        if (KEEP_HETATM[mid] < key) lo = mid + 1; else hi = mid - 1;
    }
    return false;
}
bool is_water3(const char *r) { return !strncmp(r, "HOH", 3) || !strncmp(r, "WAT", 3) || !strncmp(r, "TIP", 3); }

int aa_index(const char *n)             // set_aa_desc's switch (descriptors.c:453-507), as the shim's shim_aa_index
{
    const char a = n[0], b = n[0] ? n[1] : 0, c = (n[0] && n[1]) ? n[2] : 0;
    switch (a) {
    case 'A': return b == 'L' ? 0 : b == 'R' ? 1 : (b == 'S' && c == 'P') ? 3 : 2;
    case 'C': return 4;
    case 'G': return c == 'U' ? 6 : c == 'Y' ? 7 : 5;
    case 'H': return 8;
    case 'I': return 9;
    case 'L': return b == 'Y' ? 11 : 10;
    case 'M': return 12;
    case 'P': return b == 'H' ? 13 : 14;
    case 'S': return 15;
    case 'T': return b == 'H' ? 16 : b == 'R' ? 17 : 18;
    case 'V': return 19;
    }
    return -1;
}

}  // namespace

struct fpkio_file {
    std::string path;
    std::vector<AtomRec> watoms;        // pdb_w_lig->latoms (keep_lig = 1): a superset of pdb->latoms on the restated route
    std::vector<int32_t> nolig;         // pdb->latoms[k] = watoms[nolig[k]]
    int nhet[2] = {0, 0};
    std::vector<float> f, wf;           // 6 x n: x y z radius electroneg bfactor;  5 x nw: x y z radius electroneg
    std::vector<int32_t> iv;            // 2 x n: id res_id
    std::vector<int8_t> aa;
    std::vector<uint8_t> fl, wfl;
    int nw_count = -1;                  // -a: watoms[0..nw_count) is pdb_w_lig, the rest holds pdb's own atoms (the two lists are then independent)
    bool have_stats = false; float st_avg = 0, st_min = 0, st_max = 0;      // -a: B-factor statistics as the reference's loop over the stored atoms leaves them
    std::vector<float> xlig;            // -r: xyz triplets of the explicit ligand's records (pdb->xlig_x/y/z)
    std::vector<int32_t> lig_idx, lig_mol, name_key;   // dpocket: the known ligand (indices into watoms), its molecules, atom names of pdb packed
    fpk_structure st;
    long long bytes = 0;
};

namespace {

// rpdb_extract_pdb_atom (rpdb.c:262-352) on the persistent line buffer L
void extract_atom(const char *L, int rlen, float x, float y, float z, AtomRec &a)
{
    memset(&a, 0, sizeof a);
    slice_strncpy(a.type, L, 6);
    a.id = slice_atoi(L + 6, 5);
    slice_strncpy(a.name, L + 12, 4);
    str_trim(a.name);
    a.aloc = L[16];
    slice_strncpy(a.res_name, L + 17, 4);
    str_trim(a.res_name);
    a.chain[0] = L[21];
    a.res_id = slice_atoi(L + 22, 4);
    a.insert = L[26];
    a.x = x; a.y = y; a.z = z;                        // the same slices parsed by the caller (rpdb_extract_atom_values, rpdb.c:632-648)
    a.occ = slice_atof(L + 54, 6); a.bf = slice_atof(L + 60, 6);
    if (rlen >= 77) {
        slice_strncpy(a.symbol, L + 76, 2);
        str_trim(a.symbol);
        if (!a.symbol[0]) guess_element(a.name, a.symbol, a.res_name);
    } else guess_element(a.name, a.symbol, a.res_name);
    str_trim(a.symbol);
    if (rlen >= 79) {
        if ((L[78] == ' ' && L[79] == ' ') || L[78] == '\n') a.charge = 0;
        else { const char b[3] = {L[78], L[79], 0}; a.charge = slice_atoi(b, 2); }
    } else a.charge = 0;
    const ElemRow *e = elem_lookup(a.symbol);
    a.radius = e ? e->rvdw : -1.0f;
    a.en = e ? e->en : -1.0f;
}

bool chain_passes(const fpkio_read_opts *o, char c)     // chains_to_delete (rpdb.c:1561-1591)
{
    if (!o || o->n_chains <= 0) return !(o && o->chains_are_kept);
    bool hit = false;
    for (int i = 0; i < o->n_chains && i < 20; i++) if (o->chains[i][0] == c && o->chains[i][1] == 0) hit = true;
    return o->chains_are_kept ? hit : !hit;
}


// ---- -a (chain as ligand): the reference's two passes, literally ----------------------------------------------------------------------
// rpdb_open counts with the CURRENT record's chain, rpdb_read decides its branch with the chain of the PREVIOUS record (chainbuf is updated
// inside the branches, rpdb.c:1079-1081,1225-1227), so at every boundary between a ligand chain and another one the two disagree by one
// record: the first record of a ligand chain is stored as a protein atom, the first record after it goes down the HETATM route. The atom
// array has the size the first pass counted; stores beyond it are invisible, slots never stored stay zero (calloc). Reproduced as is.
struct PassResult { std::vector<AtomRec> atoms; int natoms = 0, nhet = 0; std::vector<float> xlig; float avg = 0, mn = 0, mx = 0; };

bool is_lig_chain(const fpkio_read_opts *o, char c)
{
    for (int i = 0; i < o->n_lig_chains && i < 20; i++) if (o->lig_chains[i][0] == c && o->lig_chains[i][1] == 0) return true;
    return false;
}

struct LineReader {             // fgets(buf, 82, f) over a memory image, with the reference's never-cleared buffer
    const char *p, *end;
    char L[LINE_MAX_CHARS + 3];
    LineReader(const char *d, size_t n) : p(d), end(d + n) { memset(L, 0, sizeof L); }
    bool next()
    {
        if (p >= end) return false;
        size_t n = (size_t)(end - p) < (size_t)LINE_MAX_CHARS ? (size_t)(end - p) : (size_t)LINE_MAX_CHARS;
        if (const void *nl = memchr(p, '\n', n)) n = (size_t)((const char *)nl - p) + 1;
        memcpy(L, p, n);
        L[n] = 0;
        p += n;
        return true;
    }
};

bool emulate_rpdb(const char *data, size_t size, const fpkio_read_opts *o, const int keep_lig, PassResult &R)
{
    const bool xlig_active = true;              // -a sets xlig_resnumber = 0 (fparams.c:201), and every test is "xlig_resnumber > -1"
    auto rmatch = [&](const char *L, int resnbuf, const char *resb) {
        return o->has_xlig && o->xlig_chain && L[21] == o->xlig_chain && resnbuf == o->xlig_resnumber && o->xlig_resname[0] == resb[0] &&
               o->xlig_resname[1] == resb[1] && o->xlig_resname[2] == resb[2];
    };
    auto resname = [](const char *L, char *resb) { slice_strncpy(resb, L + 17, 4); str_trim(resb); };
    bool model_flag = false;
    if (o->model_number > 0) { LineReader s(data, size); while (s.next()) if (!strncmp(s.L, "MODEL", 5)) { model_flag = true; break; } }
    // ---- rpdb_open: the counts (rpdb.c:830-978) ----
    int natoms = 0, nhet = 0, nxl = 0;
    {
        LineReader f(data, size);
        float x = 0, y = 0, z = 0;
        int resnbuf = 0, cur_model = 0;
        bool model_read = false;
        char resb[8] = {0};
        while (f.next()) {
            const char *L = f.L;
            if (model_flag && !strncmp(L, "MODEL", 5) && ++cur_model == o->model_number) model_read = true;
            if (model_flag && !model_read) continue;
            const bool is_atom = !strncmp(L, "ATOM ", 5), is_het = !strncmp(L, "HETATM", 6), isl = is_lig_chain(o, L[21]), ctd = chain_passes(o, L[21]);
            if (is_atom && !isl) {
                if (ctd) { resname(L, resb); x = slice_atof(L + 30, 8); y = slice_atof(L + 38, 8); z = slice_atof(L + 46, 8); resnbuf = slice_atoi(L + 22, 4); }
                if ((L[16] == ' ' || L[16] == 'A') && x < 9990 && y < 9990 && z < 9990 && ctd) natoms++;
                if (x < 9990 && y < 9990 && z < 9990 && xlig_active && rmatch(L, resnbuf, resb)) nxl++;
            } else if (is_het || (is_atom && isl)) {
                if (ctd) { x = slice_atof(L + 30, 8); y = slice_atof(L + 38, 8); z = slice_atof(L + 46, 8); resnbuf = slice_atoi(L + 22, 4); }
                if ((L[16] == ' ' || L[16] == 'A') && x < 9990 && y < 9990 && z < 9990) {
                    if (ctd) {
                        resname(L, resb);
                        if ((keep_lig && !is_water3(resb)) || (keep_lig && isl)) { natoms++; nhet++; }
                        else if (!isl && keep_hetatm(resb)) { nhet++; natoms++; }
                    }
                    if (isl) nxl++;
                }
                if (x < 9990 && y < 9990 && z < 9990) {
                    resname(L, resb);
                    resnbuf = slice_atoi(L + 22, 4);
                    if (xlig_active && ((isl && !is_water3(resb)) || rmatch(L, resnbuf, resb))) nxl++;
                }
            } else if (model_read && !strncmp(L, "ENDMDL", 6)) model_read = false;
            else if (!model_flag && !strncmp(L, "END", 3)) break;
        }
    }
    if (natoms == 0) return false;
    R.natoms = natoms; R.nhet = nhet;
    R.atoms.assign((size_t)natoms, AtomRec());
    for (auto &a : R.atoms) memset(&a, 0, sizeof a);
    // ---- rpdb_read: the stores (rpdb.c:1064-1424) ----
    std::vector<AtomRec> extra;                 // stored beyond the counted size: invisible to the rest, but inside the B-factor loop
    int iatoms = 0;
    {
        LineReader f(data, size);
        float tx = 0, ty = 0, tz = 0;
        int resnbuf = 0, cur_model = 0;
        bool model_read = false;
        char resb[8] = {0}, chainbuf = 0;       // the reference's chainbuf[0] is uninitialised before the first record
        auto store = [&](const char *L) {
            AtomRec a;
            extract_atom(L, (int)strlen(L), slice_atof(L + 30, 8), slice_atof(L + 38, 8), slice_atof(L + 46, 8), a);
            if (iatoms < natoms) R.atoms[(size_t)iatoms] = a; else extra.push_back(a);
            iatoms++;
        };
        auto push_xlig = [&](const char *L) { R.xlig.push_back(slice_atof(L + 30, 8)); R.xlig.push_back(slice_atof(L + 38, 8)); R.xlig.push_back(slice_atof(L + 46, 8)); };
        while (f.next()) {
            const char *L = f.L;
            if (model_flag && !strncmp(L, "MODEL", 5) && ++cur_model == o->model_number) model_read = true;
            if (model_flag && !model_read) continue;
            const bool is_atom = !strncmp(L, "ATOM ", 5), is_het = !strncmp(L, "HETATM", 6), prev_isl = is_lig_chain(o, chainbuf);
            if (is_atom && !prev_isl) {
                chainbuf = L[21];
                const bool ctd = chain_passes(o, chainbuf), isl = is_lig_chain(o, chainbuf);
                if (ctd) { tx = slice_atof(L + 30, 8); ty = slice_atof(L + 38, 8); tz = slice_atof(L + 46, 8); }
                if ((L[16] == ' ' || L[16] == 'A') && tx < 9990 && ty < 9990 && tz < 9990) {
                    if (ctd) { resname(L, resb); resnbuf = slice_atoi(L + 22, 4); }
                    if (isl) push_xlig(L);
                    if (ctd) store(L);          // "a simple atom": also the first record of a ligand chain
                }
                if (tx < 9990 && ty < 9990 && tz < 9990 && nxl) {
                    resname(L, resb);
                    resnbuf = slice_atoi(L + 22, 4);
                    if ((isl && !is_water3(resb)) || rmatch(L, resnbuf, resb)) push_xlig(L);
                }
            } else if (is_het || (is_atom && prev_isl)) {
                chainbuf = L[21];
                const bool ctd = chain_passes(o, chainbuf), isl = is_lig_chain(o, chainbuf);
                if (ctd) { tx = slice_atof(L + 30, 8); ty = slice_atof(L + 38, 8); tz = slice_atof(L + 46, 8); resnbuf = slice_atoi(L + 22, 4); }
                if ((L[16] == ' ' || L[16] == 'A') && tx < 9990 && ty < 9990 && tz < 9990) {
                    if (ctd) resname(L, resb);
                    if (isl) push_xlig(L);
                    if (ctd && nhet > 0) {      // "else if (pdb->lhetatm)"
                        if ((keep_lig && !is_water3(resb)) || (keep_lig && isl)) store(L);
                        else if (!isl && keep_hetatm(resb)) store(L);
                    }
                }
                if (tx < 9990 && ty < 9990 && tz < 9990 && nxl) {
                    resnbuf = slice_atoi(L + 22, 4);
                    resname(L, resb);
                    if ((isl && !is_water3(resb)) || rmatch(L, resnbuf, resb)) push_xlig(L);
                }
            } else if (model_read && !strncmp(L, "ENDMDL", 6)) model_read = false;
            else if (!model_flag && !strncmp(L, "END", 3)) break;
        }
    }
    // B-factor statistics over the STORED atoms (rpdb.c:1429-1454), hydrogens counted over the declared ones
    float avg = 0.0f, mn = 0.0f, mx = 0.0f;
    for (int i = 0; i < iatoms; i++) {
        const AtomRec &a = i < natoms ? R.atoms[(size_t)i] : extra[(size_t)(i - natoms)];
        avg += a.bf;
        if (a.bf < mn) mn = a.bf;
        if (a.bf > mx) mx = a.bf;
    }
    int nh = 0;
    for (const AtomRec &a : R.atoms) if (!strcmp(a.symbol, "H")) nh++;
    R.avg = avg / (iatoms - nh); R.mn = mn; R.mx = mx;
    if ((int)(R.xlig.size() / 3) > nxl) R.xlig.resize((size_t)3 * nxl);
    return true;
}

int parse_pdb(const char *path, const char *data, size_t size, const fpkio_read_opts *opts, fpkio_file **out)
{
    fpkio_file *F = new fpkio_file();
    F->path = path ? path : "";
    F->bytes = (long long)size;
    F->watoms.reserve(size / 81 + 8);
    char L[LINE_MAX_CHARS + 3];
    memset(L, 0, sizeof L);
    const char *p = data, *end = data + size;
    const bool chainlig = opts && opts->n_lig_chains > 0;
    if (chainlig) {
        if (opts->ligand[0]) { delete F; return FPKIO_ERR_UNSUPPORTED; }
        PassResult A, W;
        if (!emulate_rpdb(data, size, opts, 0, A) || !emulate_rpdb(data, size, opts, 1, W)) { delete F; return FPKIO_ERR_NO_ATOMS; }
        F->watoms = W.atoms;
        F->nw_count = (int)W.atoms.size();
        for (size_t k = 0; k < A.atoms.size(); k++) { F->watoms.push_back(A.atoms[k]); F->nolig.push_back((int32_t)(F->nw_count + (int)k)); }
        F->nhet[0] = A.nhet; F->nhet[1] = W.nhet;
        F->xlig = A.xlig;
        F->have_stats = true; F->st_avg = A.avg; F->st_min = A.mn; F->st_max = A.mx;
        p = end;                                // the single-pass loop below is skipped
    }
    float x = 0, y = 0, z = 0;
    const bool ligmode = opts && opts->ligand[0] && opts->ligand[1];
    const bool xlig = opts && opts->has_xlig;
    // -l N: a first pass looks for MODEL records (rpdb.c:817-828); without any, the option has no effect
    bool model_flag = false, model_read = false;
    int cur_model = 0;
    if (opts && opts->model_number > 0)
        for (const char *q = data; q < end;) {
            size_t n = (size_t)(end - q) < (size_t)LINE_MAX_CHARS ? (size_t)(end - q) : (size_t)LINE_MAX_CHARS;
            if (const void *nl = memchr(q, '\n', n)) n = (size_t)((const char *)nl - q) + 1;
            if (n >= 5 && !strncmp(q, "MODEL", 5)) { model_flag = true; break; }
            q += n;
        }
    while (p < end) {
        size_t n = (size_t)(end - p) < (size_t)LINE_MAX_CHARS ? (size_t)(end - p) : (size_t)LINE_MAX_CHARS;
        if (const void *nl = memchr(p, '\n', n)) n = (size_t)((const char *)nl - p) + 1;
        memcpy(L, p, n);
        L[n] = 0;
        p += n;
        if (model_flag && !strncmp(L, "MODEL", 5) && ++cur_model == opts->model_number) model_read = true;      // rpdb.c:834-839
        if (model_flag && !model_read) continue;
        const bool is_atom = !strncmp(L, "ATOM ", 5), is_het = !is_atom && !strncmp(L, "HETATM", 6);
        if (is_atom || is_het) {
            if (!chain_passes(opts, L[21])) continue;
            x = slice_atof(L + 30, 8); y = slice_atof(L + 38, 8); z = slice_atof(L + 46, 8);
            if (xlig && x < 9990 && y < 9990 && z < 9990) {      // rpdb.c:893-903,962-973: every record of the residue, whatever its alt-loc
                char rb[8];
                slice_strncpy(rb, L + 17, 4);
                str_trim(rb);
                if (opts->xlig_chain && L[21] == opts->xlig_chain && slice_atoi(L + 22, 4) == opts->xlig_resnumber &&
                    opts->xlig_resname[0] == rb[0] && opts->xlig_resname[1] == rb[1] && opts->xlig_resname[2] == rb[2]) {
                    F->xlig.push_back(x); F->xlig.push_back(y); F->xlig.push_back(z);
                }
            }
            if (!((L[16] == ' ' || L[16] == 'A') && x < 9990 && y < 9990 && z < 9990)) continue;
            bool in_lig = true, in_nolig = true, is_ligand_atom = false;
            char resb[8];
            slice_strncpy(resb, L + 17, 4);
            str_trim(resb);
            // dpocket: rpdb_open / rpdb_read with ligan = the ligand's residue name (rpdb.c:872-885,925-936,1131-1160,1250-1273)
            const bool lig_match = ligmode && resb[0] == opts->ligand[0] && resb[1] == opts->ligand[1] && resb[2] == opts->ligand[2];
            if (is_het) {
                const bool listed = keep_hetatm(resb);
                in_nolig = listed;                               // rpdb.c:947-961 with keep_lig = 0 (also for the ligand's own name)
                if (ligmode) { is_ligand_atom = lig_match; in_lig = lig_match || listed; }      // keep_lig = 1, ligan != NULL: nothing else is kept
                else in_lig = !is_water3(resb) || listed;        // rpdb.c:941-946 with keep_lig = 1, ligan = NULL
                if (!in_lig) continue;
            } else if (lig_match) { is_ligand_atom = true; in_nolig = false; }      // an ATOM record of the ligand: only in pdb_w_lig
            F->watoms.emplace_back();
            extract_atom(L, (int)strlen(L), x, y, z, F->watoms.back());
            if (is_het && !is_ligand_atom) F->nhet[1]++;
            if (is_het && in_nolig) F->nhet[0]++;
            if (is_ligand_atom) F->lig_idx.push_back((int32_t)F->watoms.size() - 1);
            if (in_nolig) F->nolig.push_back((int32_t)F->watoms.size() - 1);
        } else if (model_read && !strncmp(L, "ENDMDL", 6)) model_read = false;        // rpdb.c:972-973: the wanted model is over, nothing else is read
        else if (!model_flag && !strncmp(L, "END", 3)) break;    // rpdb.c:975, :1420 (model_flag == 0): ENDMDL ends the file too
    }
    if (F->nolig.empty()) { delete F; return FPKIO_ERR_NO_ATOMS; }
    // ---- the flat arrays of fpk_structure (shim_flatten / shim_job_prepare) + the s_pdb B-factor statistics (rpdb.c:1429-1454) ----
    const int n = (int)F->nolig.size(), nw = F->nw_count >= 0 ? F->nw_count : (int)F->watoms.size();
    F->f.resize((size_t)6 * n); F->iv.resize((size_t)2 * n); F->aa.resize(n); F->fl.resize(n);
    F->wf.resize((size_t)5 * nw); F->wfl.resize(nw);
    float avg = 0.0f, mn = 0.0f, mx = 0.0f;
    int nh = 0;
    for (int k = 0; k < n; k++) {
        const AtomRec &a = F->watoms[F->nolig[k]];
        F->f[k] = a.x; F->f[(size_t)n + k] = a.y; F->f[(size_t)2 * n + k] = a.z;
        F->f[(size_t)3 * n + k] = a.radius; F->f[(size_t)4 * n + k] = a.en; F->f[(size_t)5 * n + k] = a.bf;
        F->iv[k] = a.id; F->iv[(size_t)n + k] = a.res_id;
        F->aa[k] = (int8_t)aa_index(a.res_name);
        uint8_t fl = 0;
        if (!strcmp(a.symbol, "H")) { fl |= FPK_ATOM_H; nh++; }
        if (a.symbol[0] == 'H') fl |= FPK_ATOM_H1;
        if (opts && opts->n_xpocket > 0)                       // atom_in_explicit_pocket (voronoi.c:934-949)
            for (int q = 0; q < opts->n_xpocket && q < 64; q++)
                if (opts->xp_res[q] == a.res_id && opts->xp_chain[q] == a.chain[0] &&
                    (opts->xp_ins[q] == a.insert || (opts->xp_ins[q] == '-' && (a.insert == ' ' || a.insert == 0)))) { fl |= FPK_ATOM_XPOCKET; break; }
        F->fl[k] = fl;
        avg += a.bf;
        if (a.bf < mn) mn = a.bf;
        if (a.bf > mx) mx = a.bf;
    }
    avg /= (n - nh);
    for (int k = 0; k < nw; k++) {
        const AtomRec &a = F->watoms[k];
        F->wf[k] = a.x; F->wf[(size_t)nw + k] = a.y; F->wf[(size_t)2 * nw + k] = a.z; F->wf[(size_t)3 * nw + k] = a.radius; F->wf[(size_t)4 * nw + k] = a.en;
        F->wfl[k] = a.symbol[0] == 'H' ? FPK_ATOM_H : 0;
    }
    fpk_structure &s = F->st;
    memset(&s, 0, sizeof s);
    s.natoms = n;
    s.x = F->f.data(); s.y = s.x + n; s.z = s.x + 2 * (size_t)n; s.radius = s.x + 3 * (size_t)n; s.electroneg = s.x + 4 * (size_t)n; s.bfactor = s.x + 5 * (size_t)n;
    s.atom_id = F->iv.data(); s.res_id = s.atom_id + n; s.aa_idx = F->aa.data(); s.flags = F->fl.data();
    if (F->have_stats) { avg = F->st_avg; mn = F->st_min; mx = F->st_max; }
    s.avg_bfactor = avg; s.min_bfactor = mn; s.max_bfactor = mx;
    s.natoms_w = nw;
    s.wx = F->wf.data(); s.wy = s.wx + nw; s.wz = s.wx + 2 * (size_t)nw; s.wradius = s.wx + 3 * (size_t)nw; s.welectroneg = s.wx + 4 * (size_t)nw;
    s.wflags = F->wfl.data();
    s.same_pdb = 0;                     // fpmain.c:143-144: two objects
    if (!F->xlig.empty()) { s.n_xlig = (int32_t)(F->xlig.size() / 3); s.xlig = F->xlig.data(); }
    if (ligmode) {
        if (F->lig_idx.empty()) { delete F; return FPKIO_ERR_NO_LIGAND; }
        // ligand molecules: runs of equal (chain, res_id) in file order (dpocket.c:211-222, neighbor.c:619-637)
        int mol = 0;
        for (size_t j = 0; j < F->lig_idx.size(); j++) {
            const AtomRec &a = F->watoms[F->lig_idx[j]];
            if (j > 0) { const AtomRec &q = F->watoms[F->lig_idx[j - 1]]; if (q.chain[0] != a.chain[0] || q.res_id != a.res_id) mol++; }
            F->lig_mol.push_back(mol);
        }
        F->name_key.resize(n);
        for (int k = 0; k < n; k++) {                          // atm_corsp compares atom names (atom.c:281): four bytes packed
            const char *nm = F->watoms[F->nolig[k]].name;
            unsigned char b[4] = {0, 0, 0, 0};
            for (int i = 0; i < 4 && nm[i]; i++) b[i] = (unsigned char)nm[i];
            F->name_key[k] = (int)((unsigned)b[0] | ((unsigned)b[1] << 8) | ((unsigned)b[2] << 16) | ((unsigned)b[3] << 24));
        }
        s.n_lig = (int32_t)F->lig_idx.size(); s.lig_atoms = F->lig_idx.data(); s.lig_mol = F->lig_mol.data(); s.atom_name_key = F->name_key.data();
    }
    *out = F;
    return FPKIO_OK;
}

bool read_whole(const char *path, std::string &buf)
{
    int fd = open(path, O_RDONLY);
    if (fd < 0) return false;
    struct stat st;
    if (fstat(fd, &st)) { close(fd); return false; }
    buf.resize((size_t)st.st_size);
    size_t got = 0;
    while (got < buf.size()) {
        ssize_t r = read(fd, &buf[got], buf.size() - got);
        if (r <= 0) break;
        got += (size_t)r;
    }
    close(fd);
    buf.resize(got);
    return true;
}

// ====================================================== writers =========================================================================
struct Out {                            // one output file assembled in memory, written with one write()
    std::vector<char> buf;
    size_t n = 0;
    char *room(size_t k) { if (n + k > buf.size()) buf.resize((n + k) * 2 + 4096); return buf.data() + n; }
    void clear() { n = 0; }
    const char *data() const { return buf.data(); }
    size_t size() const { return n; }
    void put(const char *t, size_t k) { memcpy(room(k), t, k); n += k; }
    void put(const char *t) { put(t, strlen(t)); }
    void put(const std::string &t) { put(t.data(), t.size()); }
    void ch(char c) { *room(1) = c; n++; }
    void pad(int k) { if (k > 0) { memset(room((size_t)k), ' ', (size_t)k); n += (size_t)k; } }
    void str_left(const char *t, int w) { const int l = (int)strlen(t); put(t, (size_t)l); pad(w - l); }       // %-Ns
    void str_right(const char *t, int w) { const int l = (int)strlen(t); pad(w - l); put(t, (size_t)l); }      // %Ns
    void reversed(const char *b, int k, int w) { char *q = room((size_t)(k > w ? k : w)); int o = 0; for (; o < w - k; o++) q[o] = ' '; while (k) q[o++] = b[--k]; n += (size_t)o; }
    void integer(long long v, int w)                                                                            // %Nd
    {
        char b[24];
        int k = 0;
        const bool neg = v < 0;
        unsigned long long u = neg ? (unsigned long long)(-v) : (unsigned long long)v;
        do { b[k++] = (char)('0' + u % 10); u /= 10; } while (u);
        if (neg) b[k++] = '-';
        reversed(b, k, w);
    }
    // %W.Pf of a float-valued double. v * 10^P is exact in double for a 24-bit mantissa and P <= 4, so rounding it to the nearest integer
    // (ties to even) is what printf's exact decimal conversion prints; everything else goes through snprintf.
    void fixed(double v, int w, int prec)
    {
        static const double P10[5] = {1.0, 10.0, 100.0, 1000.0, 10000.0};
        const double av = std::fabs(v);
        if (!(av < 1e9) || prec > 4 || (double)(float)v != v) {
            char b[64];
            const int k = snprintf(b, sizeof b, "%*.*f", w, prec, v);
            put(b, (size_t)k);
            return;
        }
        unsigned long long q = (unsigned long long)std::nearbyint(av * P10[prec]);
        char b[32];
        int k = 0;
        for (int i = 0; i < prec; i++) { b[k++] = (char)('0' + q % 10); q /= 10; }
        if (prec) b[k++] = '.';
        do { b[k++] = (char)('0' + q % 10); q /= 10; } while (q);
        if (std::signbit(v)) b[k++] = '-';
        reversed(b, k, w);
    }
    void fmt(const char *f, ...) __attribute__((format(printf, 2, 3)))
    {
        char b[512];
        va_list ap;
        va_start(ap, f);
        const int k = vsnprintf(b, sizeof b, f, ap);
        va_end(ap);
        if (k > 0) put(b, (size_t)(k < (int)sizeof b ? k : (int)sizeof b - 1));
    }
};

bool write_file(const std::string &path, const Out &data, mode_t mode, long long *bytes)
{
    int fd = open(path.c_str(), O_WRONLY | O_CREAT | O_TRUNC, mode);
    if (fd < 0) return false;
    size_t done = 0;
    while (done < data.size()) {
        ssize_t r = write(fd, data.data() + done, data.size() - done);
        if (r < 0) { if (errno == EINTR) continue; close(fd); return false; }
        done += (size_t)r;
    }
    close(fd);
    if (bytes) *bytes += (long long)data.size();
    return true;
}
bool mkdir_p(const std::string &path)
{
    std::string t = path;
    for (size_t i = 1; i < t.size(); i++)
        if (t[i] == '/') { t[i] = 0; if (mkdir(t.c_str(), 0777) && errno != EEXIST) return false; t[i] = '/'; }
    return !(mkdir(t.c_str(), 0777) && errno != EEXIST);
}

void id_field(Out &o, int id) { if (id < 100000) o.integer(id, 5); else o.fmt("%05x", id); }                 // writepdb.c:101-104
void resid_field(Out &o, int r) { if (r < 10000) o.integer(r, 4); else if (r < 65536) o.fmt("%04x", r); else o.put("****"); }

// write_pdb_atom_line (writepdb.c:78-135): the occupancy column carries abpa ? abpa_sourrounding_prob : 0, the B-factor column a0
void pdb_line(Out &o, const char *rec, int id, const char *name, char aloc, const char *res, const char *chain, int res_id, char ins,
              float x, float y, float z, float a0, int abpa, const char *sym, int charge, float prob)
{
    o.str_left(rec, 6);
    id_field(o, id);
    o.ch(' ');
    o.str_right(name, 4);
    o.ch(aloc ? aloc : ' ');
    o.str_left(res, 4);
    o.put(chain);
    resid_field(o, res_id);
    o.ch(ins ? ins : ' ');
    o.pad(3);
    o.fixed(x, 8, 3); o.fixed(y, 8, 3); o.fixed(z, 8, 3);
    o.fixed(abpa ? prob : 0.0f, 6, 2);
    o.fixed(a0, 6, 2);
    o.pad(10);
    o.str_right(sym, 2);
    if (charge == -1) o.put("  "); else o.integer(charge, 2);
    o.ch('\n');
}
void atom_line(Out &o, const AtomRec &a, const float *fx)
{
    const int abpa = fx ? (int)fx[0] : 0;
    pdb_line(o, a.type, a.id, a.name, a.aloc, a.res_name, a.chain, a.res_id, a.insert, a.x, a.y, a.z, fx ? fx[1] : 0.0f, abpa, a.symbol, a.charge,
             fx ? fx[3] : 0.0f);
}
// write_pqr_vert (voronoi.c:1536-1557) -> write_pqr_atom_line (writepdb.c:168-199); electrostatic_energy is 0 on this path
void pqr_sphere(Out &o, int i, bool apolar, int resid, const float *c)
{
    o.put("ATOM  ");
    id_field(o, i);
    o.put(apolar ? "    C STP  " : "    O STP  ");
    resid_field(o, resid);
    o.pad(4);
    o.fixed(c[0], 8, 3); o.fixed(c[1], 8, 3); o.fixed(c[2], 8, 3);
    o.pad(2); o.fixed(0.0, 6, 2); o.pad(3); o.fixed(c[3], 6, 2);
    o.ch('\n');
}


// write_mmcif_atom_line (writepdb.c:201-236): the occupancy column carries dA, blank alt-loc / insertion codes become '?'
void cif_line(Out &o, const char *rec, int id, const char *sym, const char *name, char aloc, const char *res, const char *asym, int seq, char ins,
              float x, float y, float z, float occ, int charge, int res_id, const char *chain)
{
    o.str_left(rec, 7); o.ch(' ');
    { char b[16]; snprintf(b, sizeof b, "%d", id); o.str_left(b, 6); }
    o.ch(' '); o.str_right(sym, 3); o.ch(' '); o.str_right(name, 4); o.ch(' ');
    o.ch((aloc == 0 || aloc == ' ') ? '?' : aloc);
    o.ch(' '); o.str_right(res, 4); o.ch(' '); o.str_right(asym, 6); o.ch(' ');
    o.integer(seq, 0); o.ch(' ');
    o.ch((ins == 0 || ins == ' ') ? '?' : ins);
    o.ch(' '); o.fixed(x, 8, 3); o.ch(' '); o.fixed(y, 8, 3); o.ch(' '); o.fixed(z, 8, 3); o.ch(' '); o.fixed(occ, 6, 2); o.ch(' ');
    if (charge == -1) o.put(" 0"); else o.integer(charge, 2);
    o.ch(' '); o.integer(res_id, 0); o.ch(' '); o.put(chain); o.ch('\n');
}
void cif_atom(Out &o, const AtomRec &a, const float *fx)         // label_asym_id = chain, label_seq_id = res_id for a PDB input (rpdb.c:1150-1151)
{
    cif_line(o, a.type, a.id, a.symbol, a.name, a.aloc, a.res_name, a.chain, a.res_id, a.insert, a.x, a.y, a.z, fx ? fx[2] : 0.0f, a.charge, a.res_id, a.chain);
}
const char CIF_ATOM_SITE[] =                                      // writepocket.c:39-56
    "loop_\n_atom_site.group_PDB\n_atom_site.id\n_atom_site.type_symbol\n_atom_site.label_atom_id\n_atom_site.label_alt_id\n_atom_site.label_comp_id\n"
    "_atom_site.label_asym_id\n_atom_site.label_seq_id\n_atom_site.pdbx_PDB_ins_code\n_atom_site.Cartn_x\n_atom_site.Cartn_y\n_atom_site.Cartn_z\n"
    "_atom_site.occupancy\n_atom_site.pdbx_formal_charge\n_atom_site.auth_seq_id\n_atom_site.auth_asym_id\n";

void substitute(Out &o, const char *tpl, const std::string &code)
{
    for (const char *q = tpl; *q; q++) if (*q == '\001') o.put(code); else o.ch(*q);
}

}  // namespace

extern "C" {

const char *fpkio_version(void) { return "fpocket_hostio 0.1 (PDB reader + PDB-mode writers of the -F loop)"; }

int fpkio_read_pdb_mem(const char *path, const char *data, size_t size, const fpkio_read_opts *opts, fpkio_file **out)
{
    if (!data || !out) return FPKIO_ERR_BAD_ARG;
    *out = nullptr;
    return parse_pdb(path, data, size, opts, out);
}

int fpkio_read_pdb(const char *path, const fpkio_read_opts *opts, fpkio_file **out)
{
    if (!path || !out) return FPKIO_ERR_BAD_ARG;
    *out = nullptr;
    if (strstr(path, ".cif")) return FPKIO_ERR_UNSUPPORTED;         // fpmain.c:229: mmCIF goes to read_mmcif, which is not restated
    std::string buf;
    if (!read_whole(path, buf)) {
        // fopen_pdb_check_case (utils.c:704-727): the four characters before ".pdb" in upper case, then in lower case
        std::string alt = path;
        const size_t len = alt.size();
        bool ok = false;
        if (len >= 8) {
            for (size_t i = len - 8; i <= len - 5; i++) alt[i] = (char)toupper((unsigned char)alt[i]);
            ok = read_whole(alt.c_str(), buf);
            if (!ok) { for (size_t i = len - 8; i <= len - 5; i++) alt[i] = (char)tolower((unsigned char)alt[i]); ok = read_whole(alt.c_str(), buf); }
        }
        if (!ok) return FPKIO_ERR_OPEN;
    }
    return parse_pdb(path, buf.data(), buf.size(), opts, out);
}

void fpkio_file_free(fpkio_file *f) { delete f; }
const fpk_structure *fpkio_structure(const fpkio_file *f) { return f ? &f->st : nullptr; }
int fpkio_natoms(const fpkio_file *f, int with_lig) { return !f ? 0 : with_lig ? (f->nw_count >= 0 ? f->nw_count : (int)f->watoms.size()) : (int)f->nolig.size(); }
int fpkio_nhetatm(const fpkio_file *f, int with_lig) { return !f ? 0 : f->nhet[with_lig ? 1 : 0]; }

int fpkio_atom_text(const fpkio_file *f, int with_lig, int k, char *type, char *name, char *res_name, char *chain, char *symbol,
                    int32_t *irc, char *insert_aloc)
{
    if (!f || k < 0 || k >= fpkio_natoms(f, with_lig)) return FPKIO_ERR_BAD_ARG;
    const AtomRec &a = f->watoms[with_lig ? k : f->nolig[k]];
    if (type) memcpy(type, a.type, 8);
    if (name) memcpy(name, a.name, 8);
    if (res_name) memcpy(res_name, a.res_name, 8);
    if (chain) memcpy(chain, a.chain, 4);
    if (symbol) memcpy(symbol, a.symbol, 4);
    if (irc) { irc[0] = a.id; irc[1] = a.res_id; irc[2] = a.charge; }
    if (insert_aloc) { insert_aloc[0] = a.insert; insert_aloc[1] = a.aloc; }
    return FPKIO_OK;
}

int fpkio_write_output(const fpkio_file *F, const char *pdb_path, const fpk_result *R, long long *bytes_out)
{
    return fpkio_write_output_mode(F, pdb_path, R, 'p', bytes_out);
}

int fpkio_write_output_mode(const fpkio_file *F, const char *pdb_path, const fpk_result *R, char write_mode, long long *bytes_out)
{
    if (!F || !pdb_path || !R) return FPKIO_ERR_BAD_ARG;
    if (write_mode != 'p' && write_mode != 'd' && write_mode != 'm' && write_mode != 'b') return FPKIO_ERR_UNSUPPORTED;
    const bool pdb_files = write_mode == 'p' || write_mode == 'b', cif_files = write_mode == 'm' || write_mode == 'b';
    // ---- names (fpout.c:71-80): remove_ext on the whole path, then remove_path ----
    std::string code = pdb_path, dir;
    { const size_t sl = code.rfind('/'); if (sl != std::string::npos) dir = code.substr(0, sl); }
    { const size_t dot = code.rfind('.'); if (dot != std::string::npos) code.erase(dot); }
    { const size_t sl = code.rfind('/'); if (sl != std::string::npos) code.erase(0, sl + 1); }
    const std::string out_dir = dir.empty() ? code + "_out" : dir + "/" + code + "_out";
    if (!mkdir_p(out_dir)) return FPKIO_ERR_WRITE;
    const std::string base = out_dir + "/" + code;
    const int n = (int)F->nolig.size();
    const float *fx = R->atom_fx;
    bool ok = true;
    Out o;
    // ---- write_visualization (write_visu.c:56-290): only in PDB mode ----
    if (pdb_files) {
        o.clear(); substitute(o, TPL_VMD_SH, code);   ok &= write_file(base + "_VMD.sh", o, 0777, bytes_out);
        o.clear(); substitute(o, TPL_TCL, code);      ok &= write_file(base + ".tcl", o, 0666, bytes_out);
        o.clear(); substitute(o, TPL_PYMOL_SH, code); ok &= write_file(base + "_PYMOL.sh", o, 0777, bytes_out);
        o.clear(); substitute(o, TPL_PML, code);      ok &= write_file(base + ".pml", o, 0666, bytes_out);
    }
    if (cif_files) {
        o.clear(); substitute(o, TPL_CIF_VMD_SH, code);   ok &= write_file(base + "_cif_VMD.sh", o, 0777, bytes_out);
        o.clear(); substitute(o, TPL_CIF_TCL, code);      ok &= write_file(base + "_cif.tcl", o, 0666, bytes_out);
        o.clear(); substitute(o, TPL_CIF_PYMOL_SH, code); ok &= write_file(base + "_cif_PYMOL.sh", o, 0777, bytes_out);
        o.clear(); substitute(o, TPL_CIF_PML, code);      ok &= write_file(base + "_cif.pml", o, 0666, bytes_out);
    }
    // ---- <code>_out.pdb: every atom of pdb, then the spheres pocket by pocket (writepocket.c:262-305, voronoi.c:1437-1482) ----
    if (pdb_files) {
    o.clear();
    o.room((size_t)(n + R->n_spheres) * 82 + 256);
    for (int k = 0; k < n; k++) atom_line(o, F->watoms[F->nolig[k]], fx ? fx + 4 * (size_t)k : nullptr);
    for (int r = 0; r < R->n_pockets; r++) {
        const int c = R->rank_to_cluster[r];
        int i = 0;
        for (int q = R->cl_off[c]; q < R->cl_off[c + 1]; q++) {
            const int s = R->cl_sph[q];
            const float *xyz = R->sph_xyzr + 4 * (size_t)s;
            i++;
            if (R->sph_type[s] == 0) pdb_line(o, "HETATM", i, "APOL", ' ', "STP", "C", r + 1, ' ', xyz[0], xyz[1], xyz[2], 0.0f, 0, "Ve", -1, 0.0f);
            else pdb_line(o, "HETATM", n + 1 + s, " POL", ' ', "STP", "C", r + 1, ' ', xyz[0], xyz[1], xyz[2], 0.0f, 0, "Ve", -1, 0.0f);   // s_vvertice.id as the shim numbers it
        }
    }
    ok &= write_file(base + "_out.pdb", o, 0666, bytes_out);
    }
    // ---- <code>_out.cif (writepocket.c:307-357): atoms, then every sphere as "HETATM i V APOL . STP C rank" (write_mmcif_vert with no energies) ----
    if (cif_files) {
        o.clear();
        o.room((size_t)(n + R->n_spheres) * 90 + 1024);
        o.put("data_"); o.put(code); o.put("_out\n# \n"); o.put(CIF_ATOM_SITE);
        for (int k = 0; k < n; k++) cif_atom(o, F->watoms[F->nolig[k]], fx ? fx + 4 * (size_t)k : nullptr);
        for (int r = 0; r < R->n_pockets; r++) {
            const int c = R->rank_to_cluster[r];
            int i = 0;
            for (int q = R->cl_off[c]; q < R->cl_off[c + 1]; q++) {
                const float *xyz = R->sph_xyzr + 4 * (size_t)R->cl_sph[q];
                cif_line(o, "HETATM", ++i, "V", "APOL", '.', "STP", "C", r + 1, '.', xyz[0], xyz[1], xyz[2], 0.0f, -1, r + 1, "C");
            }
        }
        o.put("# \n");
        ok &= write_file(base + "_out.cif", o, 0666, bytes_out);
    }
    // ---- <code>_info.txt (fpout.c:151-213) ----
    o.clear();
    for (int r = 0; r < R->n_pockets; r++) {
        const int c = R->rank_to_cluster[r];
        const fpk_desc &d = R->desc[c];
        o.fmt("Pocket %d :\n", r + 1);
        o.put("\tScore : \t"); o.fixed(d.score, 0, 3); o.ch('\n');
        o.put("\tDruggability Score : \t"); o.fixed(d.drug_score, 0, 3); o.ch('\n');
        o.fmt("\tNumber of Alpha Spheres : \t%d\n", R->cl_off[c + 1] - R->cl_off[c]);
        o.put("\tTotal SASA : \t"); o.fixed(d.surf_vdw14, 0, 3); o.ch('\n');
        o.put("\tPolar SASA : \t"); o.fixed(d.surf_pol_vdw14, 0, 3); o.ch('\n');
        o.put("\tApolar SASA : \t"); o.fixed(d.surf_apol_vdw14, 0, 3); o.ch('\n');
        o.put("\tVolume : \t"); o.fixed(d.volume, 0, 3); o.ch('\n');
        o.put("\tMean local hydrophobic density : \t"); o.fixed(d.mean_loc_hyd_dens, 0, 3); o.ch('\n');
        o.put("\tMean alpha sphere radius :\t"); o.fixed(d.mean_asph_ray, 0, 3); o.ch('\n');
        o.put("\tMean alp. sph. solvent access : \t"); o.fixed(d.masph_sacc, 0, 3); o.ch('\n');
        o.put("\tApolar alpha sphere proportion : \t"); o.fixed(d.apolar_asphere_prop, 0, 3); o.ch('\n');
        o.put("\tHydrophobicity score:\t"); o.fixed(d.hydrophobicity_score, 0, 3); o.ch('\n');
        o.put("\tVolume score: \t "); o.fixed(d.volume_score, 0, 3); o.ch('\n');
        o.fmt("\tPolarity score:\t %d\n", d.polarity_score);
        o.fmt("\tCharge score :\t %d\n", d.charge_score);
        o.put("\tProportion of polar atoms: \t"); o.fixed(d.prop_polar_atm, 0, 3); o.ch('\n');
        o.put("\tAlpha sphere density : \t"); o.fixed(d.as_density, 0, 3); o.ch('\n');
        o.put("\tCent. of mass - Alpha Sphere max dist: \t"); o.fixed(d.as_max_dst, 0, 3); o.ch('\n');
        o.put("\tFlexibility : \t"); o.fixed(d.flex, 0, 3); o.ch('\n');
        o.ch('\n');
    }
    ok &= write_file(base + "_info.txt", o, 0666, bytes_out);
    // ---- <code>_pockets.pqr (writepocket.c:381-416): one running index over all pockets ----
    o.clear();
    o.put("HEADER\n"
          "HEADER This is a pqr format file writen by the programm fpocket.                 \n"
          "HEADER It contains all the pocket vertices found by fpocket.                    \n");
    {
        int i = 0;
        for (int r = 0; r < R->n_pockets; r++) {
            const int c = R->rank_to_cluster[r];
            for (int q = R->cl_off[c]; q < R->cl_off[c + 1]; q++) { const int s = R->cl_sph[q]; pqr_sphere(o, ++i, R->sph_type[s] == 0, r + 1, R->sph_xyzr + 4 * (size_t)s); }
        }
    }
    o.put("TER\nEND\n");
    ok &= write_file(base + "_pockets.pqr", o, 0666, bytes_out);
    // ---- pockets/pocketK_vert.pqr, pockets/pocketK_atm.pdb (writepocket.c:475-640) ----
    const std::string pdir = out_dir + "/pockets";
    if (mkdir(pdir.c_str(), 0777) && errno != EEXIST) return FPKIO_ERR_WRITE;
    std::vector<int32_t> uniq;
    char nm[64];
    for (int r = 0; r < R->n_pockets; r++) {
        const int c = R->rank_to_cluster[r];
        const fpk_desc &d = R->desc[c];
        const int first = R->cl_off[c], last = R->cl_off[c + 1];
        o.clear();
        o.put("HEADER\n"
              "HEADER This is a pqr format file writen by the programm fpocket.                 \n"
              "HEADER It represent the voronoi vertices of a single pocket found by the         \n"
              "HEADER algorithm.                                                                \n"
              "HEADER                                                                           \n");
        o.fmt("HEADER Information about the pocket %5d:\n", r + 1);
        o.put("HEADER 0  - Pocket Score                      : "); o.fixed(d.score, 0, 4); o.ch('\n');
        o.put("HEADER 1  - Drug Score                        : "); o.fixed(d.drug_score, 0, 4); o.ch('\n');
        o.fmt("HEADER 2  - Number of V. Vertices             : %5d\n", d.nb_asph);
        o.put("HEADER 3  - Mean alpha-sphere radius          : "); o.fixed(d.mean_asph_ray, 0, 4); o.ch('\n');
        o.put("HEADER 4  - Mean alpha-sphere SA              : "); o.fixed(d.masph_sacc, 0, 4); o.ch('\n');
        o.put("HEADER 5  - Mean B-factor                     : "); o.fixed(d.flex, 0, 4); o.ch('\n');
        o.put("HEADER 6  - Hydrophobicity Score              : "); o.fixed(d.hydrophobicity_score, 0, 4); o.ch('\n');
        o.fmt("HEADER 7  - Polarity Score                    : %5d\n", d.polarity_score);
        o.put("HEADER 8  - Volume Score                      : "); o.fixed(d.volume_score, 0, 4); o.ch('\n');
        o.put("HEADER 9  - Real volume (approximation)       : "); o.fixed(d.volume, 0, 4); o.ch('\n');
        o.fmt("HEADER 10 - Charge Score                      : %5d\n", d.charge_score);
        o.put("HEADER 11 - Local hydrophobic density Score   : "); o.fixed(d.mean_loc_hyd_dens, 0, 4); o.ch('\n');
        o.fmt("HEADER 12 - Number of apolar alpha sphere     : %5d\n", d.nAlphaApol);
        o.put("HEADER 13 - Proportion of apolar alpha sphere : "); o.fixed(d.apolar_asphere_prop, 0, 4); o.ch('\n');
        {
            int i = 0;
            for (int q = first; q < last; q++) { const int s = R->cl_sph[q]; pqr_sphere(o, ++i, R->sph_type[s] == 0, r + 1, R->sph_xyzr + 4 * (size_t)s); }
        }
        o.put("TER\nEND\n");
        snprintf(nm, sizeof nm, "/pocket%d_vert.pqr", r + 1);
        ok &= write_file(pdir + nm, o, 0666, bytes_out);

        if (!pdb_files && !cif_files) continue;     // writepocket.c:492-500: pocketN_atm.* only once a write mode is decided
        // atoms contacted by the pocket's spheres, unique by atom id, first seen first (writepocket.c:604-617, is_in_lst_atm compares ids)
        uniq.clear();
        for (int q = first; q < last; q++) {
            const int32_t *ng = R->sph_neigh + 4 * (size_t)R->cl_sph[q];
            for (int j = 0; j < 4; j++) {
                const int id = F->watoms[F->nolig[ng[j]]].id;
                bool seen = false;
                for (const int32_t u : uniq) if (F->watoms[F->nolig[u]].id == id) { seen = true; break; }
                if (!seen) uniq.push_back(ng[j]);
            }
        }
        if (pdb_files) {
            o.clear();
            o.put("HEADER\n"
                  "HEADER This is a pdb format file writen by the programm fpocket.                 \n"
                  "HEADER It represents the atoms contacted by the voronoi vertices of the pocket.  \n"
                  "HEADER                                                                           \n");
            o.fmt("HEADER Information about the pocket %5d:\n", r + 1);
            o.put("HEADER 0  - Pocket Score                      : "); o.fixed(d.score, 0, 4); o.ch('\n');
            o.put("HEADER 1  - Drug Score                        : "); o.fixed(d.drug_score, 0, 4); o.ch('\n');
            o.fmt("HEADER 2  - Number of alpha spheres           : %5d\n", d.nb_asph);
            o.put("HEADER 3  - Mean alpha-sphere radius          : "); o.fixed(d.mean_asph_ray, 0, 4); o.ch('\n');
            o.put("HEADER 4  - Mean alpha-sphere Solvent Acc.    : "); o.fixed(d.masph_sacc, 0, 4); o.ch('\n');
            o.put("HEADER 5  - Mean B-factor of pocket residues  : "); o.fixed(d.flex, 0, 4); o.ch('\n');
            o.put("HEADER 6  - Hydrophobicity Score              : "); o.fixed(d.hydrophobicity_score, 0, 4); o.ch('\n');
            o.fmt("HEADER 7  - Polarity Score                    : %5d\n", d.polarity_score);
            o.put("HEADER 8  - Amino Acid based volume Score     : "); o.fixed(d.volume_score, 0, 4); o.ch('\n');
            o.put("HEADER 9  - Pocket volume (Monte Carlo)       : "); o.fixed(d.volume, 0, 4); o.ch('\n');
            o.put("HEADER 10 - Pocket volume (convex hull)       : "); o.fixed(d.convex_hull_volume, 0, 4); o.ch('\n');
            o.fmt("HEADER 11 - Charge Score                      : %5d\n", d.charge_score);
            o.put("HEADER 12 - Local hydrophobic density Score   : "); o.fixed(d.mean_loc_hyd_dens, 0, 4); o.ch('\n');
            o.fmt("HEADER 13 - Number of apolar alpha sphere     : %5d\n", d.nAlphaApol);
            o.put("HEADER 14 - Proportion of apolar alpha sphere : "); o.fixed(d.apolar_asphere_prop, 0, 4); o.ch('\n');
            for (const int32_t k : uniq) atom_line(o, F->watoms[F->nolig[k]], fx ? fx + 4 * (size_t)k : nullptr);
            o.put("TER\nEND\n");
            snprintf(nm, sizeof nm, "/pocket%d_atm.pdb", r + 1);
            ok &= write_file(pdir + nm, o, 0666, bytes_out);
        }
        if (cif_files) {                            // write_pocket_mmcif (writepocket.c:642-720)
            o.clear();
            o.fmt("data_pocket%d_atm\n# \nloop_\n_struct.pdbx_descriptor\n", r + 1);
            o.put("This is a mmcif format file writen by the programm fpocket.                 \n"
                  "It represents the atoms contacted by the voronoi vertices of the pocket.  \n"
                  "                                                                           \n");
            o.fmt("Information about the pocket %5d:\n", r + 1);
            o.put("0  - Pocket Score                      : "); o.fixed(d.score, 0, 4); o.ch('\n');
            o.put("1  - Drug Score                        : "); o.fixed(d.drug_score, 0, 4); o.ch('\n');
            o.fmt("2  - Number of alpha spheres           : %5d\n", d.nb_asph);
            o.put("3  - Mean alpha-sphere radius          : "); o.fixed(d.mean_asph_ray, 0, 4); o.ch('\n');
            o.put("4  - Mean alpha-sphere Solvent Acc.    : "); o.fixed(d.masph_sacc, 0, 4); o.ch('\n');
            o.put("5  - Mean B-factor of pocket residues  : "); o.fixed(d.flex, 0, 4); o.ch('\n');
            o.put("6  - Hydrophobicity Score              : "); o.fixed(d.hydrophobicity_score, 0, 4); o.ch('\n');
            o.fmt("7  - Polarity Score                    : %5d\n", d.polarity_score);
            o.put("8  - Amino Acid based volume Score     : "); o.fixed(d.volume_score, 0, 4); o.ch('\n');
            o.put("9  - Pocket volume (Monte Carlo)       : "); o.fixed(d.volume, 0, 4); o.ch('\n');
            o.put("10  -Pocket volume (convex hull)       : "); o.fixed(d.convex_hull_volume, 0, 4); o.ch('\n');
            o.fmt("11 - Charge Score                      : %5d\n", d.charge_score);
            o.put("12 - Local hydrophobic density Score   : "); o.fixed(d.mean_loc_hyd_dens, 0, 4); o.ch('\n');
            o.fmt("13 - Number of apolar alpha sphere     : %5d\n", d.nAlphaApol);
            o.put("14 - Proportion of apolar alpha sphere : "); o.fixed(d.apolar_asphere_prop, 0, 4); o.ch('\n');
            o.put("# \n");
            o.put(CIF_ATOM_SITE);
            for (const int32_t k : uniq) cif_atom(o, F->watoms[F->nolig[k]], fx ? fx + 4 * (size_t)k : nullptr);
            o.put("# \n");
            snprintf(nm, sizeof nm, "/pocket%d_atm.cif", r + 1);
            ok &= write_file(pdir + nm, o, 0666, bytes_out);
        }
    }
    return ok ? R->n_pockets : FPKIO_ERR_WRITE;
}

}  // extern "C"

// ====================================================== the -F loop =====================================================================
namespace {

class Pool {                            // fixed pool, two priorities: writes first (they release memory), then reads
public:
    explicit Pool(int n) { for (int i = 0; i < n; i++) th_.emplace_back([this] { run(); }); }
    ~Pool()
    {
        { std::lock_guard<std::mutex> g(m_); stop_ = true; }
        cv_.notify_all();
        for (auto &t : th_) t.join();
    }
    void submit(std::function<void()> fn, bool urgent)
    {
        { std::lock_guard<std::mutex> g(m_); (urgent ? hi_ : lo_).push_back(std::move(fn)); }
        cv_.notify_one();
    }
private:
    void run()
    {
        for (;;) {
            std::function<void()> fn;
            {
                std::unique_lock<std::mutex> g(m_);
                cv_.wait(g, [this] { return stop_ || !hi_.empty() || !lo_.empty(); });
                if (!hi_.empty()) { fn = std::move(hi_.front()); hi_.pop_front(); }
                else if (!lo_.empty()) { fn = std::move(lo_.front()); lo_.pop_front(); }
                else if (stop_) return;
            }
            if (fn) fn();
        }
    }
    std::vector<std::thread> th_;
    std::deque<std::function<void()>> hi_, lo_;
    std::mutex m_;
    std::condition_variable cv_;
    bool stop_ = false;
};

struct Group {
    int first = 0, n = 0;
    std::vector<fpkio_file *> files;
    std::vector<int> rc;
    std::vector<fpk_structure> in;
    std::vector<int> idx;
    std::vector<fpk_result> out;
    std::atomic<int> pending{0};
};

}  // namespace

namespace {
// what happens to one structure once its result is there: writes files (and, for dpocket, keeps the table rows); returns the number of pockets
// written or FPKIO_ERR_*
typedef std::function<int(int k, const fpkio_file *f, const fpk_result *r, long long *bytes)> Sink;
typedef std::function<const fpkio_read_opts *(int k)> OptsOf;

int run_pipeline(const char *const *paths, int n, const OptsOf &opts_of, const fpk_params *p, fpkio_detect_fn detect,
                 fpkio_free_fn free_result, int group, int io_threads, int quiet, fpkio_stats *stats, const Sink &sink)
{
    if (!paths || n < 0 || !p || !detect || !free_result) return FPKIO_ERR_BAD_ARG;
    if (group <= 0) group = 1024;
    if (io_threads <= 0) io_threads = (int)std::thread::hardware_concurrency();
    if (io_threads <= 0) io_threads = 4;
    const double t0 = now_s();
    fpkio_stats S;
    memset(&S, 0, sizeof S);
    S.n_files = n;
    std::mutex sm;                       // guards S, the ready queue and the in-flight count
    std::condition_variable scv;
    std::deque<Group *> ready;
    int in_flight = 0, finished = 0;
    const int ngroups = (n + group - 1) / group;
    bool detect_error = false;
    {
        Pool pool(io_threads);
        // ---- detect threads: one fpk_detect_batch per group, in order of readiness; every structure then goes to the writer pool.
        // Two of them (fpk_detect_batch is re-entrant: pooled per-call contexts): while one call drains its last chunk the other one has
        // its first chunks on the device, so the GPU does not idle at group boundaries.
        int n_det = 2, taken = 0;
        if (const char *e = getenv("FPKIO_DETECT_THREADS")) n_det = std::max(1, std::min(4, atoi(e)));
        auto det_loop = [&] {
            for (;;) {
                Group *g;
                {
                    std::unique_lock<std::mutex> lk(sm);
                    scv.wait(lk, [&] { return !ready.empty() || taken == ngroups; });
                    if (ready.empty()) return;
                    g = ready.front();
                    ready.pop_front();
                    if (++taken == ngroups) scv.notify_all();
                }
                for (int k = 0; k < g->n; k++)
                    if (g->rc[k] == FPKIO_OK) { g->in.push_back(*fpkio_structure(g->files[k])); g->idx.push_back(k); }
                    else {
                        if (!quiet) {
                            if (g->rc[k] == FPKIO_ERR_OPEN) fprintf(stderr, "! File %s does not exist\n", paths[g->first + k]);
                            else if (g->rc[k] == FPKIO_ERR_NO_ATOMS) fprintf(stderr, "! File '%s' contains no atoms...\n", paths[g->first + k]);
                            else if (g->rc[k] == FPKIO_ERR_NO_LIGAND) fprintf(stdout, "ERROR - No ligand %s found in %s.\n", opts_of(g->first + k)->ligand, paths[g->first + k]);
                            else fprintf(stderr, "! %s: not a PDB file this reader handles (mmCIF input goes through build/fpocket_cuda_batch)\n", paths[g->first + k]);
                            fprintf(stderr, "! Structure reading failed!\n");
                        }
                        std::lock_guard<std::mutex> lk(sm);
                        S.n_read_failed++;
                    }
                const int nb = (int)g->in.size();
                g->out.assign((size_t)nb, fpk_result());
                for (auto &r : g->out) memset(&r, 0, sizeof r);
                int rc = 0;
                if (nb > 0) {
                    const double td = now_s();
                    rc = detect(g->in.data(), nb, p, g->out.data());
                    const double dt = now_s() - td;
                    std::lock_guard<std::mutex> lk(sm);
                    S.detect_s += dt;
                    if (rc) { detect_error = true; S.n_detect_failed += nb; }
                }
                g->pending = nb;
                auto retire = [&, g] {
                    for (fpkio_file *f : g->files) fpkio_file_free(f);
                    delete g;
                    std::lock_guard<std::mutex> lk(sm);
                    in_flight--;
                    finished++;
                    scv.notify_all();
                };
                if (nb == 0 || rc) { if (rc) for (auto &r : g->out) free_result(&r); retire(); continue; }
                for (int q = 0; q < nb; q++)
                    pool.submit([&, g, q, retire] {
                        const int k = g->idx[q];
                        const double tw = now_s();
                        long long bytes = 0;
                        fpk_result *R = &g->out[q];
                        int w = 0;
                        // search_pocket returns a list (possibly with no pocket left, fpocket.c:225) unless the clustering tree failed (:135-139)
                        const bool have_list = (R->status == FPK_OK || R->status == FPK_ERR_NO_POCKETS);
                        if (have_list) w = sink(g->first + k, g->files[k], R, &bytes);
                        else {
                            if (!quiet && R->status == FPK_ERR_NO_SPHERES && !p->skip_clustering) fprintf(stderr, "Error in creating clustering tree, return NULL pointer...breaking up");      // with -r / -P the reference returns NULL without a word (pocket.c:77)
                            if (!quiet) printf("no pockets found\n");                            // fpmain.c:204
                        }
                        if (w < 0 && !quiet) fprintf(stderr, "! could not write the results of %s\n", paths[g->first + k]);
                        const long long natoms = fpkio_natoms(g->files[k], 0);
                        const int st = R->status;
                        free_result(R);
                        fpkio_file_free(g->files[k]);
                        g->files[k] = nullptr;
                        const double dt = now_s() - tw;
                        {
                            std::lock_guard<std::mutex> lk(sm);
                            S.write_cpu_s += dt;
                            S.bytes_out += bytes;
                            S.n_atoms += natoms;
                            if (have_list && w >= 0) { S.n_written++; S.n_pockets += w; if (!w) S.n_no_pockets++; }
                            else if (st == FPK_ERR_NO_SPHERES) S.n_no_pockets++;
                            else S.n_detect_failed++;
                        }
                        if (g->pending.fetch_sub(1) == 1) retire();
                    }, true);
            }
        };
        std::vector<std::thread> det;
        for (int t = 0; t < n_det; t++) det.emplace_back(det_loop);
        // ---- this thread: bounds the groups in flight (being read, on the GPU, being written) ----
        for (int gi = 0; gi < ngroups; gi++) {
            {
                std::unique_lock<std::mutex> lk(sm);
                scv.wait(lk, [&] { return in_flight < 2 + n_det; });
                in_flight++;
            }
            Group *g = new Group();
            g->first = gi * group;
            g->n = std::min(group, n - g->first);
            g->files.assign((size_t)g->n, nullptr);
            g->rc.assign((size_t)g->n, FPKIO_ERR_OPEN);
            g->pending = g->n;
            for (int k = 0; k < g->n; k++)
                pool.submit([&, g, k] {
                    const double tr = now_s();
                    g->rc[k] = fpkio_read_pdb(paths[g->first + k], opts_of(g->first + k), &g->files[k]);
                    const double dt = now_s() - tr;
                    {
                        std::lock_guard<std::mutex> lk(sm);
                        S.read_cpu_s += dt;
                        if (g->files[k]) S.bytes_in += g->files[k]->bytes;
                    }
                    if (g->pending.fetch_sub(1) == 1) {
                        std::lock_guard<std::mutex> lk(sm);
                        ready.push_back(g);
                        scv.notify_all();
                    }
                }, false);
        }
        for (auto &t : det) t.join();
        {
            std::unique_lock<std::mutex> lk(sm);
            scv.wait(lk, [&] { return finished == ngroups; });
        }
    }
    S.wall_s = now_s() - t0;
    if (stats) *stats = S;
    return detect_error ? FPKIO_ERR_BAD_ARG : FPKIO_OK;
}

// Philox4x32-10 (Salmon et al., SC'11), the generator of the library's Monte-Carlo volume
void philox4x32(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1, uint32_t out[4])
{
    for (int r = 0; r < 10; r++) {
        const uint64_t p0 = (uint64_t)0xD2511F53u * c0, p1 = (uint64_t)0xCD9E8D57u * c2;
        const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0, n1 = (uint32_t)p1, n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1, n3 = (uint32_t)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}
float unif31(uint32_t r31, float mn, float mx) { const float rnd = ((float)(int)r31 / (float)2147483647) * (mx - mn); return mn + rnd; }   // utils.c:113-128

// write_pocket_desc (dpocket.c:387-402): one row of a dpocket table
void dpocket_row(std::string &o, const char *fc, const char *l, const fpk_desc *d, float lv, float ovlp, float dst, float c4, float c5)
{
    static const fpk_desc zero = fpk_desc();
    if (!d) d = &zero;
    const int status = dst < 4.0 ? 1 : 0;
    const int c6 = (c4 >= 0.5 && c5 >= 0.20) ? 1 : 0;                                      // tpocket.h:38-39 M_CRIT4_VAL, M_CRIT5_VAL
    const float c6_c = ((c4 - 0.5) / (1 - 0.5)) + ((c5 - 0.20) / (1 - 0.20));
    char b[1024];
    int k = snprintf(b, sizeof b, "%s %s %6.2f %2d %6.2f %4.2f %4.2f %2d %5.2f %8.2f %10.2f %5d %4.2f %6.2f %6.2f %5.2f %4.2f %7.2f %4.2f %9.2f %7.2f %5d %5.2f %5d "
                                  "%6.2f %7.2f %5.2f %5.2f %5.2f %5.2f %5.2f %5.2f %5.2f %5.2f %5.2f %5.2f %5d",
                     fc, l, ovlp, status, dst, c4, c5, c6, c6_c, lv, d->volume, d->nb_asph, d->nas_norm, d->mean_asph_ray, d->masph_sacc, d->apolar_asphere_prop,
                     d->prop_asapol_norm, d->mean_loc_hyd_dens, d->mean_loc_hyd_dens_norm, d->hydrophobicity_score, d->volume_score, d->polarity_score,
                     d->polarity_score_norm, d->charge_score, d->flex, d->prop_polar_atm, d->as_density, d->as_density_norm, d->as_max_dst, d->as_max_dst_norm,
                     d->drug_score, d->convex_hull_volume, d->surf_pol_vdw14, d->surf_pol_vdw22, d->surf_apol_vdw14, d->surf_apol_vdw22, d->n_abpa);
    if (k < 0) k = 0;
    if (k > (int)sizeof b - 1) k = (int)sizeof b - 1;
    o.append(b, (size_t)k);
    for (int i = 0; i < 20; i++) { k = snprintf(b, sizeof b, " %3d", d->aa_compo[i]); o.append(b, (size_t)k); }
    o.push_back('\n');
}
const char *const DP_HEADER =                                                              // dpocket.h:42 + the residue names of aa.c:60-80 (dpocket.c:99-103)
    "pdb lig overlap PP-crit PP-dst crit4 crit5 crit6 crit6_continue lig_vol pock_vol nb_AS nb_AS_norm mean_as_ray mean_as_solv_acc apol_as_prop "
    "apol_as_prop_norm mean_loc_hyd_dens mean_loc_hyd_dens_norm hydrophobicity_score volume_score polarity_score polarity_score_norm charge_score flex "
    "prop_polar_atm as_density as_density_norm as_max_dst as_max_dst_norm drug_score convex_hull_volume surf_pol_vdw14 surf_pol_vdw22 surf_apol_vdw14 "
    "surf_apol_vdw22 n_abpa ALA ARG ASN ASP CYS GLN GLU GLY HIS ILE LEU LYS MET PHE PRO SER THR TRP TYR VAL\n";
}  // namespace

extern "C" {

int fpkio_run_list(const char *const *paths, int n, const fpkio_read_opts *ropts, const fpk_params *p, fpkio_detect_fn detect,
                   fpkio_free_fn free_result, int group, int io_threads, int quiet, fpkio_stats *stats)
{
    if (ropts && ropts->ligand[0]) return FPKIO_ERR_UNSUPPORTED;          // a known ligand is dpocket's business: fpkio_run_dpocket
    return run_pipeline(paths, n, [ropts](int) { return ropts; }, p, detect, free_result, group, io_threads, quiet, stats,
                        [paths, ropts](int k, const fpkio_file *f, const fpk_result *r, long long *bytes) {
                            return fpkio_write_output_mode(f, paths[k], r, (ropts && ropts->write_mode) ? ropts->write_mode : 'p', bytes);
                        });
}

int fpkio_nligand(const fpkio_file *f) { return f ? (int)f->lig_idx.size() : 0; }

int fpkio_parse_custom_ligand(const char *arg, fpkio_read_opts *o)          // fparams.c:241-262: strtok(optarg, ":"): number, name, chain
{
    if (!arg || !o) return FPKIO_ERR_BAD_ARG;
    char tmp[256];
    snprintf(tmp, sizeof tmp, "%s", arg);
    int i = 0;
    o->has_xlig = 1; o->xlig_resnumber = -1; o->xlig_resname[0] = 0; o->xlig_chain = 0;
    for (char *t = strtok(tmp, ":"); t; t = strtok(nullptr, ":")) {
        i++;
        if (i == 1) o->xlig_resnumber = atoi(t);
        else if (i == 2) snprintf(o->xlig_resname, sizeof o->xlig_resname, "%s", t);
        else if (i == 3) o->xlig_chain = t[0];
    }
    if (o->xlig_resnumber <= -1) o->has_xlig = 0;          // "xlig_resnumber > -1" is the reference's test everywhere
    return FPKIO_OK;
}

int fpkio_parse_custom_pocket(const char *arg, fpkio_read_opts *o)          // fparams.c:263-300
{
    if (!arg || !o) return FPKIO_ERR_BAD_ARG;
    char tmp[512], res[256];
    snprintf(tmp, sizeof tmp, "%s", arg);
    o->n_xpocket = 0;
    char *rest = tmp;
    for (char *pt = strtok_r(rest, ".", &rest); pt && o->n_xpocket < 64; pt = strtok_r(rest, ".", &rest)) {
        const int q = o->n_xpocket++;
        o->xp_res[q] = 0; o->xp_ins[q] = 0; o->xp_chain[q] = 0;
        snprintf(res, sizeof res, "%s", pt);
        char *rest2 = res;
        int apti = 0;
        for (char *apt = strtok_r(rest2, ":", &rest2); apt; apt = strtok_r(rest2, ":", &rest2)) {
            switch (apti) {                                 // no break in the reference's switch: every field also writes the later ones
            case 0: o->xp_res[q] = (uint16_t)atoi(apt);     /* fall through */
            case 1: o->xp_ins[q] = apt[0];                  /* fall through */
            case 2: o->xp_chain[q] = apt[0];
            }
            apti++;
        }
    }
    return FPKIO_OK;
}

float fpkio_ligand_volume(const fpkio_file *f, int niter, uint64_t seed)
{
    if (!f || f->lig_idx.empty() || niter <= 0) return 0.0f;
    float xmin = 0, xmax = 0, ymin = 0, ymax = 0, zmin = 0, zmax = 0, t;
    for (size_t i = 0; i < f->lig_idx.size(); i++) {                  // atom.c:171-190, with its else-if updates
        const AtomRec &a = f->watoms[f->lig_idx[i]];
        if (i == 0) { xmin = a.x - a.radius; xmax = a.x + a.radius; ymin = a.y - a.radius; ymax = a.y + a.radius; zmin = a.z - a.radius; zmax = a.z + a.radius; }
        else {
            if (xmin > (t = a.x - a.radius)) xmin = t; else if (xmax < (t = a.x + a.radius)) xmax = t;
            if (ymin > (t = a.y - a.radius)) ymin = t; else if (ymax < (t = a.y + a.radius)) ymax = t;
            if (zmin > (t = a.z - a.radius)) zmin = t; else if (zmax < (t = a.z + a.radius)) zmax = t;
        }
    }
    const float vbox = (xmax - xmin) * (ymax - ymin) * (zmax - zmin);
    int nb_in = 0;
    for (int i = 0; i < niter; i++) {
        uint32_t o[4];
        philox4x32((uint32_t)i, 0x11Au, 0u, 0u, (uint32_t)seed, (uint32_t)(seed >> 32), o);
        const float xr = unif31(o[0] >> 1, xmin, xmax), yr = unif31(o[1] >> 1, ymin, ymax), zr = unif31(o[2] >> 1, zmin, zmax);
        for (size_t j = 0; j < f->lig_idx.size(); j++) {
            const AtomRec &a = f->watoms[f->lig_idx[j]];
            const float xt = a.x - xr, yt = a.y - yr, zt = a.z - zr;
            if ((a.radius * a.radius) > (xt * xt + yt * yt + zt * zt)) { nb_in++; break; }
        }
    }
    return ((float)nb_in) / ((float)niter) * vbox;
}

int fpkio_dpocket_rows(const fpkio_file *f, const char *fc, const char *lig, const fpk_result *r, float lig_vol, char *rows[3])
{
    if (!f || !fc || !lig || !r || !rows) return FPKIO_ERR_BAD_ARG;
    std::string t[3];
    dpocket_row(t[0], fc, lig, r->xdesc, lig_vol, 100.0f, 0.0f, 1.0f, 1.0f);               // dpocket.c:245: the explicit pocket
    for (int q = 0; q < r->n_pockets; q++) {                                              // dpocket.c:247-287
        const float ovlp = r->pk_ovlp ? r->pk_ovlp[q] : 0.0f, dst = r->pk_dst ? r->pk_dst[q] : 0.0f, c4 = r->pk_c4 ? r->pk_c4[q] : 0.0f,
                    c5 = r->pk_c5 ? r->pk_c5[q] : 0.0f;
        dpocket_row(t[(ovlp > 40.0 || dst < 4.0) ? 1 : 2], fc, lig, &r->desc[r->rank_to_cluster[q]], lig_vol, ovlp, dst, c4, c5);
    }
    for (int k = 0; k < 3; k++) { rows[k] = (char *)malloc(t[k].size() + 1); if (!rows[k]) return FPKIO_ERR_BAD_ARG; memcpy(rows[k], t[k].c_str(), t[k].size() + 1); }
    return r->n_pockets;
}
void fpkio_free_rows(char *rows[3]) { if (rows) for (int k = 0; k < 3; k++) { free(rows[k]); rows[k] = nullptr; } }

int fpkio_run_dpocket(const char *const *paths, const char *const *ligands, int n, const fpk_params *p, fpkio_detect_fn detect,
                      fpkio_free_fn free_result, const char *const out_paths[3], char write_mode, int group, int io_threads, int quiet, fpkio_stats *stats)
{
    if (!paths || !ligands || n < 0 || !p || !out_paths) return FPKIO_ERR_BAD_ARG;
    if (write_mode && write_mode != 'p' && write_mode != 'd') return FPKIO_ERR_UNSUPPORTED;
    std::vector<fpkio_read_opts> ro((size_t)n);
    for (int k = 0; k < n; k++) { memset(&ro[k], 0, sizeof ro[k]); snprintf(ro[k].ligand, sizeof ro[k].ligand, "%s", ligands[k] ? ligands[k] : ""); if (!ro[k].ligand[0] || !ro[k].ligand[1]) return FPKIO_ERR_BAD_ARG; }
    std::vector<std::string> rows((size_t)3 * n);                    // kept per complex: the tables are written in list order at the end
    const uint64_t seed = p->mc_seed;
    const int niter = p->nb_mcv_iter;
    const char wmode = write_mode ? write_mode : 'd';
    const int rc = run_pipeline(paths, n, [&ro](int k) { return &ro[k]; }, p, detect, free_result, group, io_threads, quiet, stats,
                                [&](int k, const fpkio_file *f, const fpk_result *r, long long *bytes) {
                                    const int w = fpkio_write_output_mode(f, paths[k], r, wmode, bytes);      // dpocket.c:233
                                    if (w < 0) return w;
                                    char *t[3] = {nullptr, nullptr, nullptr};
                                    const float vol = r->lig_volume;          // get_mol_volume_ptr (atom.c:156): computed with the ligand queries (K6)
                                    (void)seed; (void)niter;
                                    if (fpkio_dpocket_rows(f, paths[k], ligands[k], r, vol, t) >= 0)
                                        for (int q = 0; q < 3; q++) rows[(size_t)3 * k + q] = t[q];
                                    fpkio_free_rows(t);
                                    return w;
                                });
    long long bytes = 0;
    for (int q = 0; q < 3; q++) {
        Out o;
        o.put(DP_HEADER);
        for (int k = 0; k < n; k++) o.put(rows[(size_t)3 * k + q]);
        if (!write_file(out_paths[q], o, 0666, &bytes)) return FPKIO_ERR_WRITE;
    }
    if (stats) stats->bytes_out += bytes;
    return rc;
}

}  // extern "C"

// ====================================================== mdpocket =========================================================================
struct fpkio_dcd {
    FILE *f = nullptr;
    bool swap = false, cell = false, dim4 = false;
    int natoms = 0, nframes = 0, read = 0;
    std::vector<float> tmp;
};

namespace {
uint32_t bswap32(uint32_t v) { return (v >> 24) | ((v >> 8) & 0xff00u) | ((v << 8) & 0xff0000u) | (v << 24); }
bool dcd_int(fpkio_dcd *d, int32_t *v) { uint32_t u; if (fread(&u, 4, 1, d->f) != 1) return false; if (d->swap) u = bswap32(u); *v = (int32_t)u; return true; }
bool dcd_skip_record(fpkio_dcd *d)      // a Fortran record: length, payload, length
{
    int32_t n, m;
    if (!dcd_int(d, &n) || n < 0 || fseek(d->f, n, SEEK_CUR) || !dcd_int(d, &m) || m != n) return false;
    return true;
}
}  // namespace

extern "C" {

int fpkio_dcd_open(const char *path, fpkio_dcd **out, int *natoms, int *nframes)
{
    if (!path || !out) return FPKIO_ERR_BAD_ARG;
    *out = nullptr;
    fpkio_dcd *d = new fpkio_dcd();
    d->f = fopen(path, "rb");
    if (!d->f) { delete d; return FPKIO_ERR_OPEN; }
    uint32_t first;
    char magic[4];
    if (fread(&first, 4, 1, d->f) != 1) { fpkio_dcd_close(d); return FPKIO_ERR_UNSUPPORTED; }
    if (first == 84) d->swap = false; else if (bswap32(first) == 84) d->swap = true; else { fpkio_dcd_close(d); return FPKIO_ERR_UNSUPPORTED; }
    int32_t ic[20];
    if (fread(magic, 1, 4, d->f) != 4 || memcmp(magic, "CORD", 4) || fread(ic, 4, 20, d->f) != 20) { fpkio_dcd_close(d); return FPKIO_ERR_UNSUPPORTED; }
    if (d->swap) for (int i = 0; i < 20; i++) ic[i] = (int32_t)bswap32((uint32_t)ic[i]);
    int32_t tail;
    if (!dcd_int(d, &tail) || tail != 84) { fpkio_dcd_close(d); return FPKIO_ERR_UNSUPPORTED; }
    d->nframes = ic[0];
    const bool charmm = ic[19] != 0;
    d->cell = charmm && ic[10] != 0;
    d->dim4 = charmm && ic[11] != 0;
    if (ic[8] != 0) { fpkio_dcd_close(d); return FPKIO_ERR_UNSUPPORTED; }       // fixed atoms: frames after the first carry only the free ones
    int32_t n, m;
    if (!dcd_skip_record(d)) { fpkio_dcd_close(d); return FPKIO_ERR_UNSUPPORTED; }                     // title
    if (!dcd_int(d, &n) || n != 4 || !dcd_int(d, &d->natoms) || !dcd_int(d, &m) || m != 4 || d->natoms <= 0) { fpkio_dcd_close(d); return FPKIO_ERR_UNSUPPORTED; }
    d->tmp.resize((size_t)3 * d->natoms);
    if (natoms) *natoms = d->natoms;
    if (nframes) *nframes = d->nframes;
    *out = d;
    return FPKIO_OK;
}

int fpkio_dcd_read(fpkio_dcd *d, float *xyz, int max_frames)
{
    if (!d || !xyz || max_frames < 0) return FPKIO_ERR_BAD_ARG;
    int got = 0;
    const size_t na = (size_t)d->natoms;
    while (got < max_frames) {
        int32_t n;
        if (d->cell) { if (!dcd_skip_record(d)) break; }
        bool ok = true;
        for (int c = 0; c < 3 && ok; c++) {
            int32_t m;
            ok = dcd_int(d, &n) && n == (int32_t)(4 * na) && fread(d->tmp.data() + c * na, 4, na, d->f) == na && dcd_int(d, &m) && m == n;
        }
        if (!ok) break;                      // end of file (or a truncated frame: what was read so far stands)
        if (d->dim4 && !dcd_skip_record(d)) break;
        float *dst = xyz + (size_t)3 * na * got;
        const float *X = d->tmp.data(), *Y = X + na, *Z = Y + na;
        if (d->swap) {
            for (size_t j = 0; j < na; j++) {
                uint32_t a, b, c2;
                memcpy(&a, X + j, 4); memcpy(&b, Y + j, 4); memcpy(&c2, Z + j, 4);
                a = bswap32(a); b = bswap32(b); c2 = bswap32(c2);
                memcpy(dst + 3 * j, &a, 4); memcpy(dst + 3 * j + 1, &b, 4); memcpy(dst + 3 * j + 2, &c2, 4);
            }
        } else
            for (size_t j = 0; j < na; j++) { dst[3 * j] = X[j]; dst[3 * j + 1] = Y[j]; dst[3 * j + 2] = Z[j]; }
        got++;
        d->read++;
    }
    return got;
}

void fpkio_dcd_close(fpkio_dcd *d) { if (d) { if (d->f) fclose(d->f); delete d; } }

int fpkio_md_topology(const fpkio_file *f, fpk_structure *st)
{
    if (!f || !st) return FPKIO_ERR_BAD_ARG;
    if (fpkio_natoms(f, 0) != fpkio_natoms(f, 1) || f->nw_count >= 0) return FPKIO_ERR_UNSUPPORTED;
    *st = f->st;
    st->natoms_w = 0; st->wx = st->wy = st->wz = st->wradius = st->welectroneg = nullptr; st->wflags = nullptr;
    st->same_pdb = 1;                            // search_pocket(pdb, params, pdb): asa.c:315 compares pointers
    return FPKIO_OK;
}

int fpkio_md_write_outputs(const fpkio_file *F, const float *dens, const float *freq, const float origin[3], const int dims[3], int n_snapshots,
                           const fpk_params *p, int use_drug_score, const char *const paths[5], long long *bytes_out)
{
    if (!F || !dens || !freq || !origin || !dims || !p || !paths || n_snapshots <= 0) return FPKIO_ERR_BAD_ARG;
    const int nx = dims[0], ny = dims[1], nz = dims[2];
    const size_t ncell = (size_t)nx * ny * nz;
    const float res = 1.0f;                      // M_MDP_GRID_RESOLUTION
    std::vector<float> g[2];
    for (int q = 0; q < 2; q++) {                // normalize_grid (mdpbase.c:354-364)
        g[q].assign(q == 0 ? freq : dens, (q == 0 ? freq : dens) + ncell);
        for (float &v : g[q]) v /= (float)n_snapshots;
    }
    bool ok = true;
    Out o;
    // ---- project_grid_on_atoms(densgrid) + write_first_bfactor_density (mdpbase.c:301-340, mdpout.c:177-191) ----
    {
        const float ng1 = (float)2.0 / res;      // M_MDP_ATOM_DENSITY_DIST / resolution
        const std::vector<float> &D = g[1];
        const int n = (int)F->nolig.size();
        o.clear();
        for (int k = 0; k < n; k++) {
            const AtomRec &a = F->watoms[F->nolig[k]];
            float density = 0.0f;
            int ix = (int)floor(((double)a.x - 2.0 - (double)origin[0]) / (double)res);
            const float maxix = (float)(int)ceil(((double)a.x + 2.0 - (double)origin[0]) / (double)res);
            const float maxiy = (float)(int)ceil(((double)a.y + 2.0 - (double)origin[1]) / (double)res);
            const float maxiz = (float)(int)ceil(((double)a.z + 2.0 - (double)origin[2]) / (double)res);
            while (ix <= maxix && nx > ix && ix >= 0) {
                int iy = (int)floor(((double)a.y - 2.0 - (double)origin[1]) / (double)res);
                while (iy <= maxiy && ny > iy && iy >= 0) {
                    int iz = (int)floor(((double)a.z - 2.0 - (double)origin[2]) / (double)res);
                    while (iz <= maxiz && nz > iz && iz >= 0) { density += D[((size_t)ix * ny + iy) * nz + iz]; iz++; }
                    iy++;
                }
                ix++;
            }
            const float bf = (float)log((double)(1 + density / (ng1) / (float)n_snapshots));
            pdb_line(o, "ATOM", a.id, a.name, a.aloc, a.res_name, a.chain, a.res_id, a.insert, a.x, a.y, a.z, bf, 0, a.symbol, a.charge, 0.0f);
        }
        o.put("TER\nEND\n");
        ok &= write_file(paths[4], o, 0666, bytes_out);
    }
    // ---- write_md_grid (mdpout.c:71-122): DX file + iso-surface PDB, frequency grid then density grid ----
    const float iso[2] = {0.5f, 8.0f};           // M_MDP_DEFAULT_ISO_VALUE_FREQ / _DENS
    for (int q = 0; q < 2; q++) {
        Out dx, pi;
        dx.room(ncell * 7 + 1024);
        dx.put("# Data calculated by mdpocket, part of the fpocket package\n"
               "# This is a standard DX file of occurences of cavities within MD trajectories.\n"
               "# The file can be visualised using the freely available VMD software\n"
               "# fpocket parameters used to create this dx file : \n");
        dx.fmt("# \t-m %2.f (min alpha sphere size) -M %.2f (max alpha sphere size)\n", p->asph_min_size, p->asph_max_size);
        dx.fmt("# \t-i %d (min number of alpha spheres per pocket)\n", p->min_pock_nb_asph);
        if (use_drug_score) dx.put("# \t-S (Map drug score to density map!)\n");
        dx.fmt("object 1 class gridpositions counts %d %d %d\n", nx, ny, nz);
        dx.fmt("origin %.2f %.2f %.2f\n", origin[0], origin[1], origin[2]);
        dx.fmt("delta %.2f 0 0\ndelta 0 %.2f 0\ndelta 0 0 %.2f\n", res, res, res);
        dx.fmt("object 2 class gridconnections counts %d %d %d\n", nx, ny, nz);
        dx.fmt("object 3 class array type double rank 0 items %d data follows\n", nx * ny * nz);
        int i = 0;
        size_t cnt = 0, idx = 0;
        for (int cx = 0; cx < nx; cx++)
            for (int cy = 0; cy < ny; cy++)
                for (int cz = 0; cz < nz; cz++, idx++) {
                    if (i == 3) { i = 0; dx.ch('\n'); }
                    const float cv = g[q][idx];
                    dx.fixed(cv, 0, 3); dx.ch(' ');
                    if (cv >= iso[q]) {
                        cnt++;
                        const float rx = origin[0] + cx * res, ry = origin[1] + cy * res, rz = origin[2] + cz * res;
                        pi.put("ATOM  "); pi.integer((long long)(int)cnt, 5); pi.put("  C   PTH     1    ");
                        pi.fixed(rx, 8, 3); pi.fixed(ry, 8, 3); pi.fixed(rz, 8, 3); pi.fixed(0.0, 6, 2); pi.fixed(0.0, 6, 2); pi.ch('\n');
                    }
                    i++;
                }
        ok &= write_file(paths[q], dx, 0666, bytes_out);
        ok &= write_file(paths[2 + q], pi, 0666, bytes_out);
    }
    return ok ? FPKIO_OK : FPKIO_ERR_WRITE;
}

}  // extern "C"
