// Host-side mirror of the reference's chip interface, recording instead of computing.
//
// The reference's Circuit::synthesize (benches/bench.rs:132-225) calls RSAChip ->
// BigIntChip -> maingate on a RegionCtx, one row at a time, with the witness values
// computed on the CPU inside every call.  The sequence of calls is data independent
// (Stack 2), so this mirror runs the SAME call sequence ONCE, symbolically:
// every MainGate / RangeChip primitive records
//   * which advice cell (column, row) it assigns,
//   * a value node saying how that cell's value derives from earlier values,
//   * the fixed-column coefficients and copy constraints of the row (for keygen).
// The recorded program is then replayed on the GPU for a whole batch of instances
// (witness.cu).  Names, argument order and panics/errors follow the reference:
//   MainGate / RangeChip   <- maingate crate (halo2wrong rev 63bde545; third party)
//   BigIntChip             <- src/big_integer/chip.rs, trait src/big_integer/instructions.rs
//   RSAChip                <- src/chip.rs, trait src/instructions.rs
#pragma once
#include <stdint.h>

#include <algorithm>
#include <array>
#include <map>
#include <stdexcept>
#include <string>
#include <vector>

namespace b2r {
namespace circuit {

// 256-bit unsigned constants (canonical), enough for word_max (134 bits) and field constants
struct U256 {
    uint64_t l[4] = {0, 0, 0, 0};
    U256() {}
    explicit U256(uint64_t v) { l[0] = v; }
    bool operator<(const U256& o) const {
        for (int i = 3; i >= 0; i--)
            if (l[i] != o.l[i]) return l[i] < o.l[i];
        return false;
    }
    bool operator==(const U256& o) const { return l[0] == o.l[0] && l[1] == o.l[1] && l[2] == o.l[2] && l[3] == o.l[3]; }
    bool is_zero() const { return (l[0] | l[1] | l[2] | l[3]) == 0; }
    static U256 pow2(unsigned b) {
        U256 r;
        r.l[b >> 6] = 1ull << (b & 63);
        return r;
    }
    U256 add(const U256& o) const {
        U256 r;
        unsigned __int128 c = 0;
        for (int i = 0; i < 4; i++) {
            c += (unsigned __int128)l[i] + o.l[i];
            r.l[i] = (uint64_t)c;
            c >>= 64;
        }
        return r;
    }
    U256 sub(const U256& o) const {
        U256 r;
        unsigned __int128 br = 0;
        for (int i = 0; i < 4; i++) {
            unsigned __int128 t = (unsigned __int128)l[i] - o.l[i] - (uint64_t)br;
            r.l[i] = (uint64_t)t;
            br = (t >> 64) & 1;
        }
        return r;
    }
    U256 mul(const U256& o) const {  // low 256 bits
        uint64_t t[8] = {0};
        for (int i = 0; i < 4; i++) {
            unsigned __int128 c = 0;
            for (int j = 0; j + i < 4; j++) {
                c += (unsigned __int128)l[i] * o.l[j] + t[i + j];
                t[i + j] = (uint64_t)c;
                c >>= 64;
            }
        }
        U256 r;
        for (int i = 0; i < 4; i++) r.l[i] = t[i];
        return r;
    }
    unsigned bits() const {
        for (int i = 3; i >= 0; i--)
            if (l[i]) return 64 * i + (64 - __builtin_clzll(l[i]));
        return 0;
    }
    U256 shr(unsigned b) const {
        U256 r;
        const unsigned w = b >> 6, s = b & 63;
        for (unsigned i = 0; i + w < 4; i++) {
            r.l[i] = l[i + w] >> s;
            if (s && i + w + 1 < 4) r.l[i] |= l[i + w + 1] << (64 - s);
        }
        return r;
    }
    U256 low(unsigned b) const {  // value mod 2^b
        U256 r = *this;
        for (unsigned i = 0; i < 4; i++) {
            if (64 * i >= b) r.l[i] = 0;
            else if (64 * (i + 1) > b) r.l[i] &= (~0ull) >> (64 * (i + 1) - b);
        }
        return r;
    }
    int pow2_log() const {  // b if value == 2^b, else -1
        unsigned b = bits();
        if (b == 0) return -1;
        return (*this == pow2(b - 1)) ? (int)(b - 1) : -1;
    }
};
// BN254 Fr modulus, for "-1" style coefficients
inline U256 fr_modulus() {
    U256 r;
    r.l[0] = 0x43e1f593f0000001ull;
    r.l[1] = 0x2833e84879b97091ull;
    r.l[2] = 0xb85045b68181585dull;
    r.l[3] = 0x30644e72e131a029ull;
    return r;
}
inline U256 fr_neg(const U256& a) { return a.is_zero() ? a : fr_modulus().sub(a); }

// ---- value nodes --------------------------------------------------------------------------
enum Op : uint8_t {
    OP_CONST = 0,   // a = constant index
    OP_INPUT,       // a = input word index (n limbs | sig limbs | hash limbs), a 64-bit integer
    OP_ADD,         // a + b
    OP_SUB,         // a - b
    OP_MUL,         // a * b
    OP_MULADD,      // a * b + c
    OP_ADDC,        // a + const[b]
    OP_ADD2C,       // a + b + const[c]
    OP_NOT,         // 1 - a
    OP_SELECT,      // c == 1 ? a : b
    OP_ISZERO,      // a == 0 ? 1 : 0
    OP_INVORONE,    // a == 0 ? 1 : a^-1
    OP_SHR,         // int(a) >> b
    OP_LOWBITS,     // int(a) mod 2^b
    OP_SUBLIMB,     // (int(a) >> b) mod 2^c
    OP_CLEARLOW,    // int(a) with its low b bits cleared
    OP_BIG,         // a = big-op index; writes its outputs to the following value ids
    OP_BIGOUT,      // written by the preceding OP_BIG
};
struct Node {
    uint8_t op;
    uint8_t pad[3];
    uint32_t a, b, c;
};
enum BigKind : uint32_t { BIG_MULMOD = 0, BIG_SUB = 1 };
struct BigOp {
    uint32_t kind;
    uint32_t limb_width;
    uint32_t na, nb, nn;   // limb counts of the operands (nn = 0 for BIG_SUB)
    uint32_t in_off;       // offset into big_inputs: a ids, b ids, n ids
    uint32_t nout;         // MULMOD: nb quotient limbs then na remainder limbs; SUB: na limbs
    uint32_t pad;
};

struct AssignedValue {
    int col = -1;
    uint32_t row = 0;
    int32_t vid = -1;  // value node
};
using AssignedCondition = AssignedValue;

struct SynthError : std::runtime_error {
    int code;
    SynthError(int c, const std::string& m) : std::runtime_error(m), code(c) {}
};

static constexpr int NUM_ADVICE = 5;
enum FixedCol { F_SA = 0, F_SB, F_SC, F_SD, F_SE, F_MUL_AB, F_MUL_CD, F_SE_NEXT, F_CONST, NUM_FIXED };

// RegionCtx + the recording state (regions are stacked by SimpleFloorPlanner: one running offset)
class RegionCtx {
  public:
    std::vector<Node> nodes;
    std::vector<uint32_t> level;
    std::vector<U256> constants;
    std::map<U256, uint32_t> const_index;
    std::vector<BigOp> big_ops;
    std::vector<uint32_t> big_inputs;
    std::vector<int32_t> cell[NUM_ADVICE];             // value id per row, -1 = unassigned (zero)
    std::vector<std::array<uint32_t, NUM_FIXED>> fixed;  // constant index per fixed column per row
    std::vector<std::array<uint8_t, 4>> range_tags;      // s_comp, tag_comp, s_over, tag_over
    std::vector<std::array<uint32_t, 4>> copies;         // (col, row, col, row)
    uint32_t offset = 0;
    uint32_t max_rows;
    int tag_of_bits[80] = {0};

    explicit RegionCtx(uint32_t max_rows_) : max_rows(max_rows_) {
        const_id(U256(0));
        const_id(U256(1));
    }
    uint32_t const_id(const U256& v) {
        auto it = const_index.find(v);
        if (it != const_index.end()) return it->second;
        uint32_t id = (uint32_t)constants.size();
        constants.push_back(v);
        const_index[v] = id;
        return id;
    }
    int32_t node(Op op, uint32_t a = 0, uint32_t b = 0, uint32_t c = 0, uint32_t lvl = 0) {
        Node n;
        n.op = op;
        n.pad[0] = n.pad[1] = n.pad[2] = 0;
        n.a = a;
        n.b = b;
        n.c = c;
        nodes.push_back(n);
        level.push_back(lvl);
        return (int32_t)nodes.size() - 1;
    }
    uint32_t lvl(int32_t v) const { return level[v]; }
    int32_t constant(const U256& v) { return node(OP_CONST, const_id(v)); }
    int32_t input(uint32_t word) { return node(OP_INPUT, word); }
    int32_t op1(Op op, int32_t a, uint32_t b = 0, uint32_t c = 0) { return node(op, a, b, c, lvl(a) + 1); }
    int32_t op2(Op op, int32_t a, int32_t b, uint32_t c = 0) { return node(op, a, b, c, std::max(lvl(a), lvl(b)) + 1); }
    int32_t op3(Op op, int32_t a, int32_t b, int32_t c) {
        return node(op, a, b, c, std::max(std::max(lvl(a), lvl(b)), lvl(c)) + 1);
    }
    // value of a recorded constant node, if it is one
    bool const_value(int32_t v, U256* out) const {
        if (nodes[v].op != OP_CONST) return false;
        *out = constants[nodes[v].a];
        return true;
    }
    void ensure_row(uint32_t row) {
        if (row >= max_rows) throw SynthError(-5, "circuit does not fit the available rows");
        if (fixed.size() <= row) {
            std::array<uint32_t, NUM_FIXED> z;
            z.fill(0);
            fixed.resize(row + 1, z);
            range_tags.resize(row + 1, std::array<uint8_t, 4>{0, 0, 0, 0});
            for (int i = 0; i < NUM_ADVICE; i++) cell[i].resize(row + 1, -1);
        }
    }
    void constrain_equal(const AssignedValue& a, const AssignedValue& b) {
        copies.push_back({(uint32_t)a.col, a.row, (uint32_t)b.col, b.row});
    }
    void next() { offset++; }
};

// ---- maingate ------------------------------------------------------------------------------
struct Term {
    enum Kind { Zero, Assigned, Unassigned } kind = Zero;
    AssignedValue src;   // Assigned
    int32_t vid = -1;    // value placed in the cell
    U256 base;           // fixed coefficient
    static Term zero() { return Term(); }
    static Term assigned(const AssignedValue& a, const U256& base) {
        Term t;
        t.kind = Assigned;
        t.src = a;
        t.vid = a.vid;
        t.base = base;
        return t;
    }
    static Term unassigned(int32_t vid, const U256& base) {
        Term t;
        t.kind = Unassigned;
        t.vid = vid;
        t.base = base;
        return t;
    }
    static Term assigned_to_mul(const AssignedValue& a) { return assigned(a, U256(0)); }
    static Term assigned_to_add(const AssignedValue& a) { return assigned(a, U256(1)); }
    static Term assigned_to_sub(const AssignedValue& a) { return assigned(a, fr_neg(U256(1))); }
    static Term unassigned_to_mul(int32_t v) { return unassigned(v, U256(0)); }
    static Term unassigned_to_add(int32_t v) { return unassigned(v, U256(1)); }
    static Term unassigned_to_sub(int32_t v) { return unassigned(v, fr_neg(U256(1))); }
};

struct CombinationOption {
    U256 s_mul_ab, s_mul_cd, se_next;
    static CombinationOption OneLinerAdd() { return CombinationOption(); }
    static CombinationOption OneLinerMul() {
        CombinationOption o;
        o.s_mul_ab = U256(1);
        return o;
    }
    static CombinationOption CombineToNextAdd(const U256& next) {
        CombinationOption o;
        o.se_next = next;
        return o;
    }
    static CombinationOption OneLinerDoubleMul(const U256& e) {
        CombinationOption o;
        o.s_mul_ab = U256(1);
        o.s_mul_cd = e;
        return o;
    }
};

class MainGate {
  public:
    // MainGate::apply: up to 5 terms go to columns a..e of one row
    std::array<AssignedValue, NUM_ADVICE> apply(RegionCtx& ctx, const std::vector<Term>& terms, const U256& constant,
                                                const CombinationOption& opt) const {
        uint32_t row = ctx.offset;
        ctx.ensure_row(row);
        std::array<AssignedValue, NUM_ADVICE> out;
        for (int i = 0; i < NUM_ADVICE; i++) {
            Term t = i < (int)terms.size() ? terms[i] : Term::zero();
            ctx.cell[i][row] = t.kind == Term::Zero ? -1 : t.vid;
            ctx.fixed[row][F_SA + i] = ctx.const_id(t.base);
            out[i].col = i;
            out[i].row = row;
            out[i].vid = t.kind == Term::Zero ? -1 : t.vid;
            if (t.kind == Term::Assigned) ctx.constrain_equal(t.src, out[i]);
        }
        ctx.fixed[row][F_MUL_AB] = ctx.const_id(opt.s_mul_ab);
        ctx.fixed[row][F_MUL_CD] = ctx.const_id(opt.s_mul_cd);
        ctx.fixed[row][F_SE_NEXT] = ctx.const_id(opt.se_next);
        ctx.fixed[row][F_CONST] = ctx.const_id(constant);
        ctx.next();
        return out;
    }
    AssignedValue assign_constant(RegionCtx& ctx, const U256& constant) const {
        int32_t v = ctx.constant(constant);
        return apply(ctx, {Term::unassigned_to_sub(v)}, constant, CombinationOption::OneLinerAdd())[0];
    }
    AssignedValue assign_value(RegionCtx& ctx, int32_t v) const {
        return apply(ctx, {Term::unassigned_to_mul(v)}, U256(0), CombinationOption::OneLinerAdd())[0];
    }
    AssignedCondition assign_bit(RegionCtx& ctx, int32_t bit) const {
        auto o = apply(ctx, {Term::unassigned_to_mul(bit), Term::unassigned_to_mul(bit), Term::unassigned_to_sub(bit)},
                       U256(0), CombinationOption::OneLinerMul());
        ctx.constrain_equal(o[0], o[1]);
        ctx.constrain_equal(o[1], o[2]);
        return o[2];
    }
    AssignedValue add_with_constant(RegionCtx& ctx, const AssignedValue& a, const AssignedValue& b, const U256& k) const {
        int32_t v = k.is_zero() ? ctx.op2(OP_ADD, a.vid, b.vid) : ctx.op2(OP_ADD2C, a.vid, b.vid, ctx.const_id(k));
        return apply(ctx, {Term::assigned_to_add(a), Term::assigned_to_add(b), Term::unassigned_to_sub(v)}, k,
                     CombinationOption::OneLinerAdd())[2];
    }
    AssignedValue add(RegionCtx& ctx, const AssignedValue& a, const AssignedValue& b) const {
        return add_with_constant(ctx, a, b, U256(0));
    }
    AssignedValue add_constant(RegionCtx& ctx, const AssignedValue& a, const U256& k) const {
        int32_t v = ctx.op1(OP_ADDC, a.vid, ctx.const_id(k));
        return apply(ctx, {Term::assigned_to_add(a), Term::unassigned_to_sub(v)}, k, CombinationOption::OneLinerAdd())[1];
    }
    AssignedValue sub(RegionCtx& ctx, const AssignedValue& a, const AssignedValue& b) const {
        int32_t v = ctx.op2(OP_SUB, a.vid, b.vid);
        return apply(ctx, {Term::assigned_to_add(a), Term::assigned_to_sub(b), Term::unassigned_to_sub(v)}, U256(0),
                     CombinationOption::OneLinerAdd())[2];
    }
    AssignedValue mul(RegionCtx& ctx, const AssignedValue& a, const AssignedValue& b) const {
        int32_t v = ctx.op2(OP_MUL, a.vid, b.vid);
        return apply(ctx, {Term::assigned_to_mul(a), Term::assigned_to_mul(b), Term::unassigned_to_sub(v)}, U256(0),
                     CombinationOption::OneLinerMul())[2];
    }
    AssignedValue mul_add(RegionCtx& ctx, const AssignedValue& a, const AssignedValue& b, const AssignedValue& to_add) const {
        int32_t v = ctx.op3(OP_MULADD, a.vid, b.vid, to_add.vid);
        return apply(ctx, {Term::assigned_to_mul(a), Term::assigned_to_mul(b), Term::assigned_to_add(to_add), Term::unassigned_to_sub(v)},
                     U256(0), CombinationOption::OneLinerMul())[3];
    }
    AssignedCondition and_(RegionCtx& ctx, const AssignedCondition& a, const AssignedCondition& b) const { return mul(ctx, a, b); }
    AssignedCondition not_(RegionCtx& ctx, const AssignedCondition& c) const {
        int32_t v = ctx.op1(OP_NOT, c.vid);
        return apply(ctx, {Term::assigned_to_add(c), Term::unassigned_to_add(v)}, fr_neg(U256(1)), CombinationOption::OneLinerAdd())[1];
    }
    AssignedValue select(RegionCtx& ctx, const AssignedValue& a, const AssignedValue& b, const AssignedCondition& cond) const {
        int32_t v = ctx.op3(OP_SELECT, a.vid, b.vid, cond.vid);
        auto o = apply(ctx, {Term::assigned_to_mul(cond), Term::assigned_to_mul(a), Term::assigned_to_mul(cond), Term::assigned_to_add(b),
                             Term::unassigned_to_sub(v)},
                       U256(0), CombinationOption::OneLinerDoubleMul(fr_neg(U256(1))));
        ctx.constrain_equal(o[0], o[2]);
        return o[4];
    }
    // MainGate::compose (maingate, third party): sum of the terms, laid out like decompose(): 4 terms per row in a..d,
    // the running remainder (base -1) in e, rows chained with CombineToNextAdd(1); returns the first row's e cell.
    // `total` is the value node of the whole sum; the remainders of later rows are `total` with its low bits cleared,
    // which holds because every caller composes ascending powers of two (to_bits).
    AssignedValue compose_bits(RegionCtx& ctx, const std::vector<AssignedCondition>& bits, int32_t total) const {
        const size_t n = bits.size(), nchunks = (n - 1) / 4 + 1;
        AssignedValue result;
        for (size_t ch = 0; ch < nchunks; ch++) {
            std::vector<Term> t;
            for (size_t j = 4 * ch; j < 4 * ch + 4 && j < n; j++) t.push_back(Term::assigned(bits[j], U256::pow2((unsigned)j)));
            while (t.size() < 4) t.push_back(Term::zero());
            const int32_t rem = ch == 0 ? total : ctx.op1(OP_CLEARLOW, total, (uint32_t)(4 * ch));
            t.push_back(Term::unassigned_to_sub(rem));
            const bool is_final = ch == nchunks - 1;
            auto o = apply(ctx, t, U256(0), is_final ? CombinationOption::OneLinerAdd() : CombinationOption::CombineToNextAdd(U256(1)));
            if (ch == 0) result = o[4];
        }
        return result;
    }
    // MainGate::to_bits (maingate, third party): number_of_bits boolean cells, least significant first, whose
    // composition is constrained equal to `composed`
    std::vector<AssignedCondition> to_bits(RegionCtx& ctx, const AssignedValue& composed, unsigned number_of_bits) const {
        if (number_of_bits == 0 || number_of_bits > 254) throw SynthError(-1, "to_bits: number_of_bits out of range");
        std::vector<AssignedCondition> bits;
        for (unsigned i = 0; i < number_of_bits; i++) bits.push_back(assign_bit(ctx, ctx.op1(OP_SUBLIMB, composed.vid, i, 1)));
        const int32_t total = ctx.op1(OP_LOWBITS, composed.vid, number_of_bits);  // what the bits compose to
        AssignedValue result = compose_bits(ctx, bits, total);
        assert_equal(ctx, result, composed);
        return bits;
    }
    // invert(): r bit, then (a * a') - 1 + r = 0 and r * a' - r = 0
    AssignedCondition is_zero(RegionCtx& ctx, const AssignedValue& a) const {
        int32_t rv = ctx.op1(OP_ISZERO, a.vid);
        int32_t iv = ctx.op1(OP_INVORONE, a.vid);
        AssignedCondition r = assign_bit(ctx, rv);
        AssignedValue a_inv = apply(ctx, {Term::assigned_to_mul(a), Term::unassigned_to_mul(iv), Term::assigned_to_add(r)},
                                    fr_neg(U256(1)), CombinationOption::OneLinerMul())[1];
        apply(ctx, {Term::assigned_to_mul(r), Term::assigned_to_mul(a_inv), Term::assigned_to_sub(r)}, U256(0),
              CombinationOption::OneLinerMul());
        return r;
    }
    AssignedCondition is_equal(RegionCtx& ctx, const AssignedValue& a, const AssignedValue& b) const {
        AssignedValue d = sub(ctx, a, b);
        return is_zero(ctx, d);
    }
    void assert_equal(RegionCtx& ctx, const AssignedValue& a, const AssignedValue& b) const { ctx.constrain_equal(a, b); }
    void assert_equal_to_constant(RegionCtx& ctx, const AssignedValue& a, const U256& k) const {
        apply(ctx, {Term::assigned_to_add(a)}, fr_neg(k), CombinationOption::OneLinerAdd());
    }
    void assert_zero(RegionCtx& ctx, const AssignedValue& a) const { assert_equal_to_constant(ctx, a, U256(0)); }
    void assert_one(RegionCtx& ctx, const AssignedValue& a) const { assert_equal_to_constant(ctx, a, U256(1)); }
};

class RangeChip {
  public:
    // RangeChip::assign(ctx, value, limb_bit_len, bit_len): decompose into sublimbs, 4 per row
    // in a..d with the running remainder in e; lookups enabled on every row.
    AssignedValue assign(RegionCtx& ctx, int32_t value, unsigned limb_bit_len, unsigned bit_len) const {
        MainGate mg;
        unsigned nl = bit_len / limb_bit_len, over = bit_len % limb_bit_len;
        if (over) nl++;
        if (!ctx.tag_of_bits[limb_bit_len] || (over && !ctx.tag_of_bits[over]))
            throw SynthError(-6, "range table for this bit length was not configured");
        unsigned nchunks = (nl - 1) / 4 + 1;
        AssignedValue result;
        for (unsigned ch = 0; ch < nchunks; ch++) {
            std::vector<Term> t;
            for (unsigned j = 4 * ch; j < 4 * ch + 4 && j < nl; j++) {
                int32_t sv = ctx.op1(OP_SUBLIMB, value, j * limb_bit_len, limb_bit_len);
                t.push_back(Term::unassigned(sv, U256::pow2(j * limb_bit_len)));
            }
            while (t.size() < 4) t.push_back(Term::zero());
            int32_t rem = ch == 0 ? value : ctx.op1(OP_CLEARLOW, value, 4 * ch * limb_bit_len);
            t.push_back(Term::unassigned_to_sub(rem));
            bool is_final = ch == nchunks - 1;
            uint32_t row = ctx.offset;
            ctx.ensure_row(row);
            ctx.range_tags[row][0] = 1;
            ctx.range_tags[row][1] = (uint8_t)ctx.tag_of_bits[limb_bit_len];
            if (is_final && over) {
                ctx.range_tags[row][2] = 1;
                ctx.range_tags[row][3] = (uint8_t)ctx.tag_of_bits[over];
            }
            auto o = mg.apply(ctx, t, U256(0), is_final ? CombinationOption::OneLinerAdd() : CombinationOption::CombineToNextAdd(U256(1)));
            if (ch == 0) result = o[4];
        }
        return result;
    }
};

// ---- BigIntChip (src/big_integer/chip.rs) -------------------------------------------------------
struct AssignedInteger {
    std::vector<AssignedValue> limbs;  // little-endian limbs
    size_t num_limbs() const { return limbs.size(); }
    const AssignedValue& limb(size_t i) const { return limbs[i]; }
    void extend_limbs(size_t n, const AssignedValue& zero) {
        for (size_t i = 0; i < n; i++) limbs.push_back(zero);
    }
};
struct UnassignedInteger {
    std::vector<int32_t> limbs;  // value ids
};

class BigIntChip {
  public:
    unsigned limb_width, num_limbs;
    static constexpr unsigned NUM_LOOKUP_LIMBS = 8;
    MainGate main_gate_;
    RangeChip range_chip_;

    // chip.rs:1174-1186
    BigIntChip(unsigned limb_width_, unsigned bits_len) : limb_width(limb_width_) {
        if (bits_len % limb_width != 0) throw SynthError(-1, "bits_len % limb_width != 0");
        num_limbs = bits_len / limb_width;
        if (compute_mul_word_max(limb_width, num_limbs).bits() > 254) throw SynthError(-1, "mul word max exceeds the field");
    }
    const MainGate& main_gate() const { return main_gate_; }
    const RangeChip& range_chip() const { return range_chip_; }
    static unsigned sublimb_bit_len(unsigned b) {  // chip.rs:1357-1365
        unsigned v = b / NUM_LOOKUP_LIMBS;
        return v == 0 ? 1 : v;
    }
    static U256 compute_mul_word_max(unsigned limb_width, unsigned min_n) {  // chip.rs:1368-1372
        U256 m = U256::pow2(limb_width).sub(U256(1));
        return U256(min_n).mul(m).mul(m).add(m);
    }
    // chip.rs:1220-1249: (composition_bit_lens, overflow_bit_lens)
    static void compute_range_lens(unsigned limb_width, unsigned num_limbs, std::vector<unsigned>& comp, std::vector<unsigned>& over) {
        unsigned out_comp = limb_width / NUM_LOOKUP_LIMBS, out_over = limb_width % out_comp;
        unsigned fresh_bits = U256::pow2(limb_width).add(U256::pow2(limb_width)).bits() - limb_width;
        unsigned fresh_comp = sublimb_bit_len(fresh_bits), fresh_over = fresh_bits % fresh_comp;
        U256 wm = compute_mul_word_max(limb_width, num_limbs);
        unsigned mul_bits = wm.add(wm).bits() - limb_width;
        unsigned mul_comp = sublimb_bit_len(mul_bits), mul_over = mul_bits % mul_comp;
        comp = {out_comp, fresh_comp, mul_comp};
        over = {out_over, fresh_over, mul_over};
    }

    // chip.rs:62-82
    AssignedInteger assign_integer(RegionCtx& ctx, const UnassignedInteger& integer) const {
        AssignedInteger r;
        for (int32_t limb : integer.limbs) r.limbs.push_back(range_chip_.assign(ctx, limb, sublimb_bit_len(limb_width), limb_width));
        return r;
    }
    // chip.rs:1252-1281 (integer given as 64-bit words, little endian)
    AssignedInteger assign_constant(RegionCtx& ctx, const std::vector<uint64_t>& integer_words, size_t max_num_limbs) const {
        // bits of the integer
        size_t bits = 0;
        for (size_t i = integer_words.size(); i-- > 0;)
            if (integer_words[i]) {
                bits = 64 * i + (64 - __builtin_clzll(integer_words[i]));
                break;
            }
        size_t nl = bits % limb_width == 0 ? bits / limb_width : bits / limb_width + 1;
        if (nl > max_num_limbs) throw SynthError(-6, "assign_constant: integer has more limbs than allowed");
        if (limb_width != 64) throw SynthError(-1, "assign_constant: limb_width must be 64");
        AssignedInteger r;
        for (size_t i = 0; i < nl; i++) r.limbs.push_back(main_gate_.assign_constant(ctx, U256(integer_words[i])));
        AssignedValue zero = main_gate_.assign_constant(ctx, U256(0));
        for (size_t i = nl; i < max_num_limbs; i++) r.limbs.push_back(zero);
        return r;
    }
    AssignedInteger assign_constant_fresh(RegionCtx& ctx, const std::vector<uint64_t>& w) const { return assign_constant(ctx, w, num_limbs); }
    AssignedInteger assign_constant_muled(RegionCtx& ctx, const std::vector<uint64_t>& w, size_t l, size_t r) const {
        return assign_constant(ctx, w, l + r - 1);
    }
    // chip.rs:130-147
    AssignedInteger max_value(RegionCtx& ctx, size_t n) const {
        AssignedInteger r;
        U256 limb_max = U256::pow2(limb_width).sub(U256(1));
        for (size_t i = 0; i < n; i++) r.limbs.push_back(main_gate_.assign_constant(ctx, limb_max));
        return r;
    }
    // chip.rs:245-297
    AssignedInteger add(RegionCtx& ctx, const AssignedInteger& a_in, const AssignedInteger& b_in) const {
        size_t n1 = a_in.num_limbs(), n2 = b_in.num_limbs(), max_n = n1 < n2 ? n2 : n1;
        AssignedValue zero_value = main_gate_.assign_constant(ctx, U256(0));
        AssignedInteger a = a_in, b = b_in;
        a.extend_limbs(max_n - n1, zero_value);
        b.extend_limbs(max_n - n2, zero_value);
        std::vector<AssignedValue> c_vals;
        AssignedValue carry = zero_value;
        AssignedValue limb_max_val = main_gate_.assign_constant(ctx, U256::pow2(limb_width));
        for (size_t i = 0; i < max_n; i++) {
            AssignedValue a_b = main_gate_.add(ctx, a.limb(i), b.limb(i));
            AssignedValue sum = main_gate_.add(ctx, a_b, carry);
            int32_t c_val = ctx.op1(OP_LOWBITS, sum.vid, limb_width);
            int32_t carry_val = ctx.op1(OP_SHR, sum.vid, limb_width);
            AssignedValue c = range_chip_.assign(ctx, c_val, sublimb_bit_len(limb_width), limb_width);
            AssignedValue cy = range_chip_.assign(ctx, carry_val, sublimb_bit_len(limb_width), limb_width);
            AssignedValue c_add_carry = main_gate_.mul_add(ctx, cy, limb_max_val, c);
            main_gate_.assert_equal(ctx, sum, c_add_carry);
            c_vals.push_back(c);
            carry = cy;
        }
        c_vals.push_back(carry);
        AssignedInteger r;
        r.limbs = c_vals;
        return r;
    }
    // chip.rs:1286-1318
    AssignedInteger sub_unchecked(RegionCtx& ctx, const AssignedInteger& a, const AssignedInteger& b) const {
        if (a.num_limbs() < b.num_limbs()) throw SynthError(-6, "sub_unchecked: a has fewer limbs than b");
        size_t max_n = a.num_limbs();
        int32_t first = big_op(ctx, BIG_SUB, a, b, nullptr, (uint32_t)max_n);
        AssignedInteger c;
        for (size_t i = 0; i < max_n; i++)
            c.limbs.push_back(range_chip_.assign(ctx, first + 1 + (int32_t)i, sublimb_bit_len(limb_width), limb_width));
        AssignedInteger added = add(ctx, b, c);
        assert_equal_fresh(ctx, a, added);
        return c;
    }
    // chip.rs:310-373: returns (a - b or b - a, is_overflowed)
    AssignedInteger sub(RegionCtx& ctx, const AssignedInteger& a, const AssignedInteger& b, AssignedValue* is_overflowed) const {
        size_t n2 = b.num_limbs();
        AssignedInteger max_int = max_value(ctx, n2);
        AssignedInteger inflated_a = add(ctx, a, max_int);
        AssignedInteger inflated_subed = sub_unchecked(ctx, inflated_a, b);
        AssignedValue one = main_gate_.assign_bit(ctx, ctx.constant(U256(1)));
        AssignedValue is_not_overflowed = main_gate_.is_equal(ctx, inflated_subed.limb(n2), one);
        *is_overflowed = main_gate_.not_(ctx, is_not_overflowed);
        size_t num_l = inflated_subed.num_limbs(), num_r = a.num_limbs() > n2 ? a.num_limbs() : n2;
        AssignedValue zero_value = main_gate_.assign_constant(ctx, U256(0));
        AssignedInteger sel_l, sel_r;
        for (size_t i = 0; i < num_l; i++) {
            if (i >= n2) sel_l.limbs.push_back(main_gate_.select(ctx, inflated_subed.limb(i), zero_value, is_not_overflowed));
            else sel_l.limbs.push_back(main_gate_.select(ctx, inflated_subed.limb(i), b.limb(i), is_not_overflowed));
        }
        for (size_t i = 0; i < num_r; i++) {
            if (i >= a.num_limbs()) sel_r.limbs.push_back(main_gate_.select(ctx, max_int.limb(i), zero_value, is_not_overflowed));
            else if (i >= n2) sel_r.limbs.push_back(main_gate_.select(ctx, zero_value, a.limb(i), is_not_overflowed));
            else sel_r.limbs.push_back(main_gate_.select(ctx, max_int.limb(i), a.limb(i), is_not_overflowed));
        }
        return sub_unchecked(ctx, sel_l, sel_r);
    }
    // chip.rs:386-419: unreduced convolution
    AssignedInteger mul(RegionCtx& ctx, const AssignedInteger& a, const AssignedInteger& b) const {
        size_t d0 = a.num_limbs(), d1 = b.num_limbs(), d = d0 + d1 - 1;
        AssignedInteger c;
        for (size_t i = 0; i < d; i++) {
            AssignedValue acc = main_gate_.assign_constant(ctx, U256(0));
            size_t j = d1 >= i + 1 ? 0 : i + 1 - d1;
            while (j < d0 && j <= i) {
                size_t k = i - j;
                acc = main_gate_.mul_add(ctx, a.limb(j), b.limb(k), acc);
                j++;
            }
            c.limbs.push_back(acc);
        }
        return c;
    }
    AssignedInteger square(RegionCtx& ctx, const AssignedInteger& a) const { return mul(ctx, a, a); }
    // RefreshAux::new (src/big_integer/mod.rs:431-482): how many extra limbs every limb of a product of num_limbs_l x
    // num_limbs_r maximal limbs spills into when it is cut back to limb_width bits
    static std::vector<size_t> refresh_increased_limbs(unsigned limb_width, size_t num_limbs_l, size_t num_limbs_r) {
        const U256 max_limb = U256::pow2(limb_width).sub(U256(1));
        const size_t d = num_limbs_l + num_limbs_r - 1;
        std::vector<U256> muled;
        for (size_t i = 0; i < d; i++) {
            size_t j = num_limbs_r >= i + 1 ? 0 : i + 1 - num_limbs_r;
            U256 acc;
            while (j < num_limbs_l && j <= i) {
                acc = acc.add(max_limb.mul(max_limb));
                j++;
            }
            muled.push_back(acc);
        }
        std::vector<size_t> inc;
        size_t cur_d = 0;
        const size_t max_d = d;
        while (cur_d <= max_d) {
            if (muled.size() <= cur_d) muled.push_back(U256());
            const unsigned bits = muled[cur_d].bits();
            const size_t num_chunks = bits % limb_width == 0 ? bits / limb_width : bits / limb_width + 1;
            if (num_chunks == 0) throw SynthError(-1, "RefreshAux: empty limb");  // the reference underflows here (usize)
            inc.push_back(num_chunks - 1);
            std::vector<U256> chunks;
            for (size_t c = 0; c < num_chunks; c++) {
                chunks.push_back(muled[cur_d].low(limb_width));
                muled[cur_d] = muled[cur_d].shr(limb_width);
            }
            for (size_t j = 0; j < num_chunks; j++) {
                if (muled.size() <= cur_d + j) muled.push_back(U256());
                muled[cur_d + j] = muled[cur_d + j].add(chunks[j]);
            }
            cur_d++;
        }
        return inc;
    }
    // chip.rs:168-233: Muled -> Fresh
    AssignedInteger refresh(RegionCtx& ctx, const AssignedInteger& a, size_t num_limbs_l, size_t num_limbs_r) const {
        const std::vector<size_t> inc = refresh_increased_limbs(limb_width, num_limbs_l, num_limbs_r);
        if (a.num_limbs() != num_limbs_l + num_limbs_r - 1) throw SynthError(-6, "refresh: limb count does not match the aux data");
        const size_t num_limbs_fresh = inc.size();
        AssignedValue zero_val = main_gate_.assign_constant(ctx, U256(0));
        std::vector<AssignedValue> refreshed = a.limbs;
        while (refreshed.size() < num_limbs_fresh) refreshed.push_back(zero_val);
        AssignedValue limb_max = main_gate_.assign_constant(ctx, U256::pow2(limb_width));
        for (size_t i = 0; i < num_limbs_fresh; i++) {
            AssignedValue limb = refreshed[i];
            for (size_t j = 0; j < inc[i] + 1; j++) {
                AssignedValue q, n;
                div_mod_main_gate(ctx, limb, limb_max, &q, &n);
                if (j == 0) {
                    refreshed[i] = n;
                } else {
                    if (i + j >= refreshed.size()) throw SynthError(-6, "refresh: carry past the last limb");  // index panic in the reference
                    refreshed[i + j] = main_gate_.add(ctx, refreshed[i + j], n);
                }
                limb = q;
            }
            main_gate_.assert_zero(ctx, limb);
        }
        AssignedInteger r;
        for (size_t i = 0; i < num_limbs_fresh; i++) {
            AssignedValue range_assigned = range_chip_.assign(ctx, refreshed[i].vid, sublimb_bit_len(limb_width), limb_width);
            main_gate_.assert_equal(ctx, refreshed[i], range_assigned);
            r.limbs.push_back(refreshed[i]);
        }
        return r;
    }
    // chip.rs:452-481
    AssignedInteger add_mod(RegionCtx& ctx, const AssignedInteger& a, const AssignedInteger& b, const AssignedInteger& n) const {
        AssignedInteger added = add(ctx, a, b);
        AssignedValue is_overflowed;
        AssignedInteger subed = sub(ctx, added, n, &is_overflowed);
        const size_t nl = subed.num_limbs();
        if (nl < added.num_limbs()) throw SynthError(-6, "add_mod: limb count underflow");
        AssignedValue zero_value = main_gate_.assign_constant(ctx, U256(0));
        added.extend_limbs(nl - added.num_limbs(), zero_value);
        std::vector<AssignedValue> res;
        for (size_t i = 0; i < nl; i++) res.push_back(main_gate_.select(ctx, added.limb(i), subed.limb(i), is_overflowed));
        for (size_t i = n.num_limbs(); i < nl; i++) main_gate_.assert_zero(ctx, res[i]);
        AssignedInteger r;
        r.limbs.assign(res.begin(), res.begin() + n.num_limbs());
        return r;
    }
    // chip.rs:495-529
    AssignedInteger sub_mod(RegionCtx& ctx, const AssignedInteger& a, const AssignedInteger& b, const AssignedInteger& n) const {
        AssignedValue is_overflowed1, is_overflowed2;
        AssignedInteger subed1 = sub(ctx, a, b, &is_overflowed1);
        AssignedInteger subed2 = sub(ctx, n, subed1, &is_overflowed2);
        main_gate_.assert_zero(ctx, is_overflowed2);
        const size_t nl = subed2.num_limbs();
        if (nl < subed1.num_limbs()) throw SynthError(-6, "sub_mod: limb count underflow");
        AssignedValue zero_value = main_gate_.assign_constant(ctx, U256(0));
        subed1.extend_limbs(nl - subed1.num_limbs(), zero_value);
        std::vector<AssignedValue> res;
        for (size_t i = 0; i < nl; i++) res.push_back(main_gate_.select(ctx, subed2.limb(i), subed1.limb(i), is_overflowed1));
        for (size_t i = n.num_limbs(); i < nl; i++) main_gate_.assert_zero(ctx, res[i]);
        AssignedInteger r;
        r.limbs.assign(res.begin(), res.begin() + n.num_limbs());
        return r;
    }
    // chip.rs:664-696: variable exponent, exp_limb_bits bits taken from every limb of e (least significant first)
    AssignedInteger pow_mod(RegionCtx& ctx, const AssignedInteger& a, const AssignedInteger& e, const AssignedInteger& n, unsigned exp_limb_bits) const {
        std::vector<AssignedValue> e_bits;
        for (const auto& limb : e.limbs) {
            std::vector<AssignedCondition> bits = main_gate_.to_bits(ctx, limb, exp_limb_bits);
            e_bits.insert(e_bits.end(), bits.begin(), bits.end());
        }
        AssignedInteger acc = assign_constant_fresh(ctx, {1});
        AssignedInteger squared = a;
        for (const auto& e_bit : e_bits) {
            AssignedInteger muled = mul_mod(ctx, acc, squared, n);
            for (size_t j = 0; j < acc.num_limbs(); j++) acc.limbs[j] = main_gate_.select(ctx, muled.limb(j), acc.limb(j), e_bit);
            squared = square_mod(ctx, squared, n);
        }
        return acc;
    }
    // chip.rs:542-629
    AssignedInteger mul_mod(RegionCtx& ctx, const AssignedInteger& a, const AssignedInteger& b, const AssignedInteger& n) const {
        size_t n1 = a.num_limbs(), n2 = b.num_limbs();
        if (n1 != n.num_limbs()) throw SynthError(-6, "mul_mod: a and n must have the same number of limbs");
        int32_t first = big_op(ctx, BIG_MULMOD, a, b, &n, (uint32_t)(n1 + n2));
        AssignedInteger quotient_int, prod_int;
        for (size_t i = 0; i < n2; i++)
            quotient_int.limbs.push_back(range_chip_.assign(ctx, first + 1 + (int32_t)i, sublimb_bit_len(limb_width), limb_width));
        for (size_t i = 0; i < n1; i++)
            prod_int.limbs.push_back(range_chip_.assign(ctx, first + 1 + (int32_t)(n2 + i), sublimb_bit_len(limb_width), limb_width));
        AssignedInteger ab = mul(ctx, a, b);
        AssignedInteger qn = mul(ctx, quotient_int, n);
        size_t n_sum = n1 + n2;
        AssignedInteger eq_a, eq_b;
        for (size_t i = 0; i < n_sum - 1; i++) {
            eq_a.limbs.push_back(ab.limb(i));
            if (i < n1) eq_b.limbs.push_back(main_gate_.add(ctx, qn.limb(i), prod_int.limb(i)));
            else eq_b.limbs.push_back(qn.limb(i));
        }
        assert_equal_muled(ctx, eq_a, eq_b, n1, n2);
        return prod_int;
    }
    AssignedInteger square_mod(RegionCtx& ctx, const AssignedInteger& a, const AssignedInteger& n) const { return mul_mod(ctx, a, a, n); }
    // chip.rs:710-742: e as little-endian bytes
    AssignedInteger pow_mod_fixed_exp(RegionCtx& ctx, const AssignedInteger& a, const std::vector<uint8_t>& e_le, const AssignedInteger& n) const {
        size_t num_e_bits = 0;
        for (size_t i = e_le.size() * 8; i-- > 0;)
            if ((e_le[i >> 3] >> (i & 7)) & 1) {
                num_e_bits = i + 1;
                break;
            }
        AssignedInteger acc = assign_constant(ctx, {1}, a.num_limbs());
        AssignedInteger squared = a;
        for (size_t i = 0; i < num_e_bits; i++) {
            AssignedInteger cur_sq = squared;
            squared = square_mod(ctx, cur_sq, n);
            if (!((e_le[i >> 3] >> (i & 7)) & 1)) continue;
            acc = mul_mod(ctx, acc, cur_sq, n);
        }
        return acc;
    }
    // chip.rs:754-767
    AssignedValue is_zero(RegionCtx& ctx, const AssignedInteger& a) const {
        AssignedValue bit = main_gate_.assign_bit(ctx, ctx.constant(U256(1)));
        for (const auto& limb : a.limbs) {
            AssignedValue z = main_gate_.is_zero(ctx, limb);
            bit = main_gate_.and_(ctx, bit, z);
        }
        return bit;
    }
    // chip.rs:780-805
    AssignedValue is_equal_fresh(RegionCtx& ctx, const AssignedInteger& a, const AssignedInteger& b) const {
        size_t n1 = a.num_limbs(), n2 = b.num_limbs();
        bool a_larger = n1 > n2;
        size_t max_n = a_larger ? n1 : n2;
        AssignedValue eq_bit = main_gate_.assign_bit(ctx, ctx.constant(U256(1)));
        for (size_t i = 0; i < max_n; i++) {
            AssignedValue flag;
            if (a_larger && i >= n2) flag = main_gate_.is_zero(ctx, a.limb(i));
            else if (!a_larger && i >= n1) flag = main_gate_.is_zero(ctx, b.limb(i));
            else flag = main_gate_.is_equal(ctx, a.limb(i), b.limb(i));
            eq_bit = main_gate_.and_(ctx, eq_bit, flag);
        }
        return eq_bit;
    }
    // chip.rs:1323-1349 (n must be a recorded power-of-two constant: it always is limb_max)
    void div_mod_main_gate(RegionCtx& ctx, const AssignedValue& a, const AssignedValue& n, AssignedValue* q_out, AssignedValue* m_out) const {
        U256 nv;
        int lg = ctx.const_value(n.vid, &nv) ? nv.pow2_log() : -1;
        if (lg < 0) throw SynthError(-6, "div_mod_main_gate: divisor must be a constant power of two");
        int32_t qv = ctx.op1(OP_SHR, a.vid, (uint32_t)lg);
        int32_t mv = ctx.op1(OP_LOWBITS, a.vid, (uint32_t)lg);
        AssignedValue q = main_gate_.assign_value(ctx, qv);
        AssignedValue a_mod_n = main_gate_.assign_value(ctx, mv);
        AssignedValue nq = main_gate_.mul(ctx, n, q);
        AssignedValue a_sub_nq = main_gate_.sub(ctx, a, nq);
        main_gate_.assert_equal(ctx, a_mod_n, a_sub_nq);
        *q_out = q;
        *m_out = a_mod_n;
    }
    // chip.rs:822-895
    AssignedValue is_equal_muled(RegionCtx& ctx, const AssignedInteger& a, const AssignedInteger& b, size_t num_limbs_l, size_t num_limbs_r) const {
        size_t min_n = num_limbs_r >= num_limbs_l ? num_limbs_l : num_limbs_r;
        U256 word_max = compute_mul_word_max(limb_width, (unsigned)min_n);
        size_t nl = num_limbs_l + num_limbs_r - 1;
        unsigned word_max_width = word_max.add(word_max).bits();
        unsigned carry_bits = word_max_width - limb_width;
        AssignedValue limb_max = main_gate_.assign_constant(ctx, U256::pow2(limb_width));
        AssignedValue accumulated_extra = main_gate_.assign_constant(ctx, U256(0));
        std::vector<AssignedValue> carry, cs;
        carry.push_back(main_gate_.assign_constant(ctx, U256(0)));
        AssignedValue eq_bit = main_gate_.assign_bit(ctx, ctx.constant(U256(1)));
        for (size_t i = 0; i < nl; i++) {
            AssignedValue a_b = main_gate_.sub(ctx, a.limb(i), b.limb(i));
            AssignedValue sum = main_gate_.add_with_constant(ctx, a_b, carry[i], word_max);
            AssignedValue new_carry, c;
            div_mod_main_gate(ctx, sum, limb_max, &new_carry, &c);
            carry.push_back(new_carry);
            cs.push_back(c);
            accumulated_extra = main_gate_.add_constant(ctx, accumulated_extra, word_max);
            AssignedValue q_acc, mod_acc;
            div_mod_main_gate(ctx, accumulated_extra, limb_max, &q_acc, &mod_acc);
            AssignedValue cs_acc_eq = main_gate_.is_equal(ctx, cs[i], mod_acc);
            eq_bit = main_gate_.and_(ctx, eq_bit, cs_acc_eq);
            accumulated_extra = q_acc;
            if (i < nl - 1) {
                AssignedValue range_assigned = range_chip_.assign(ctx, carry[i + 1].vid, sublimb_bit_len(carry_bits), carry_bits);
                AssignedValue range_eq = main_gate_.is_equal(ctx, carry[i + 1], range_assigned);
                eq_bit = main_gate_.and_(ctx, eq_bit, range_eq);
            } else {
                AssignedValue final_carry_eq = main_gate_.is_equal(ctx, carry[i + 1], accumulated_extra);
                eq_bit = main_gate_.and_(ctx, eq_bit, final_carry_eq);
            }
        }
        return eq_bit;
    }
    // chip.rs:908-1006
    AssignedValue is_less_than_or_equal(RegionCtx& ctx, const AssignedInteger& a, const AssignedInteger& b) const {
        AssignedValue is_overflowed;
        sub(ctx, a, b, &is_overflowed);
        return is_overflowed;
    }
    AssignedValue is_less_than(RegionCtx& ctx, const AssignedInteger& a, const AssignedInteger& b) const {
        AssignedValue is_overflowed = is_less_than_or_equal(ctx, a, b);
        AssignedValue is_eq = is_equal_fresh(ctx, a, b);
        AssignedValue is_not_eq = main_gate_.not_(ctx, is_eq);
        return main_gate_.and_(ctx, is_overflowed, is_not_eq);
    }
    AssignedValue is_greater_than(RegionCtx& ctx, const AssignedInteger& a, const AssignedInteger& b) const {
        return main_gate_.not_(ctx, is_less_than_or_equal(ctx, a, b));
    }
    AssignedValue is_greater_than_or_equal(RegionCtx& ctx, const AssignedInteger& a, const AssignedInteger& b) const {
        return main_gate_.not_(ctx, is_less_than(ctx, a, b));
    }
    AssignedValue is_in_field(RegionCtx& ctx, const AssignedInteger& a, const AssignedInteger& n) const { return is_less_than(ctx, a, n); }
    // chip.rs:1016-1158
    void assert_zero(RegionCtx& ctx, const AssignedInteger& a) const { main_gate_.assert_one(ctx, is_zero(ctx, a)); }
    void assert_equal_fresh(RegionCtx& ctx, const AssignedInteger& a, const AssignedInteger& b) const {
        main_gate_.assert_one(ctx, is_equal_fresh(ctx, a, b));
    }
    void assert_equal_muled(RegionCtx& ctx, const AssignedInteger& a, const AssignedInteger& b, size_t n1, size_t n2) const {
        main_gate_.assert_one(ctx, is_equal_muled(ctx, a, b, n1, n2));
    }
    void assert_less_than(RegionCtx& ctx, const AssignedInteger& a, const AssignedInteger& b) const { main_gate_.assert_one(ctx, is_less_than(ctx, a, b)); }
    void assert_less_than_or_equal(RegionCtx& ctx, const AssignedInteger& a, const AssignedInteger& b) const {
        main_gate_.assert_one(ctx, is_less_than_or_equal(ctx, a, b));
    }
    void assert_greater_than(RegionCtx& ctx, const AssignedInteger& a, const AssignedInteger& b) const { main_gate_.assert_one(ctx, is_greater_than(ctx, a, b)); }
    void assert_greater_than_or_equal(RegionCtx& ctx, const AssignedInteger& a, const AssignedInteger& b) const {
        main_gate_.assert_one(ctx, is_greater_than_or_equal(ctx, a, b));
    }
    void assert_in_field(RegionCtx& ctx, const AssignedInteger& a, const AssignedInteger& n) const { main_gate_.assert_one(ctx, is_in_field(ctx, a, n)); }

  private:
    // records a multi-limb BigUint computation (the only "real" bignum math of the witness,
    // chip.rs:556-567 / :1297-1300); returns the id of the OP_BIG node, outputs follow it
    int32_t big_op(RegionCtx& ctx, BigKind kind, const AssignedInteger& a, const AssignedInteger& b, const AssignedInteger* n, uint32_t nout) const {
        BigOp op;
        op.kind = kind;
        op.limb_width = limb_width;
        op.na = (uint32_t)a.num_limbs();
        op.nb = (uint32_t)b.num_limbs();
        op.nn = n ? (uint32_t)n->num_limbs() : 0;
        op.in_off = (uint32_t)ctx.big_inputs.size();
        op.nout = nout;
        op.pad = 0;
        uint32_t lv = 0;
        auto push = [&](const AssignedInteger& x) {
            for (const auto& l : x.limbs) {
                ctx.big_inputs.push_back((uint32_t)l.vid);
                lv = std::max(lv, ctx.lvl(l.vid));
            }
        };
        push(a);
        push(b);
        if (n) push(*n);
        uint32_t idx = (uint32_t)ctx.big_ops.size();
        ctx.big_ops.push_back(op);
        int32_t first = ctx.node(OP_BIG, idx, 0, 0, lv + 1);
        for (uint32_t i = 0; i < nout; i++) ctx.node(OP_BIGOUT, idx, i, 0, lv + 1);
        return first;
    }
};

// ---- RSAChip (src/chip.rs) ------------------------------------------------------------------------
struct AssignedRSAPublicKey {
    AssignedInteger n;
    std::vector<uint8_t> e_fix;  // RSAPubE::Fix, little-endian bytes
    AssignedInteger e_var;       // RSAPubE::Var (src/lib.rs:58-63): assigned exponent limbs; empty for Fix
};
struct AssignedRSASignature {
    AssignedInteger c;
};

class RSAChip {
  public:
    static constexpr unsigned LIMB_WIDTH = 64;
    unsigned bits_len, exp_limb_bits;
    RSAChip(unsigned bits_len_, unsigned exp_limb_bits_) : bits_len(bits_len_), exp_limb_bits(exp_limb_bits_) {}
    BigIntChip bigint_chip() const { return BigIntChip(LIMB_WIDTH, bits_len); }
    // chip.rs:249-254
    static void compute_range_lens(unsigned num_limbs, std::vector<unsigned>& comp, std::vector<unsigned>& over) {
        BigIntChip::compute_range_lens(LIMB_WIDTH, num_limbs, comp, over);
        comp.push_back(32 / BigIntChip::NUM_LOOKUP_LIMBS);
    }
    // chip.rs:58-70 (fixed exponent: no cells for e)
    AssignedRSAPublicKey assign_public_key(RegionCtx& ctx, const UnassignedInteger& n, const std::vector<uint8_t>& e_fix) const {
        AssignedRSAPublicKey pk;
        pk.n = bigint_chip().assign_integer(ctx, n);
        pk.e_fix = e_fix;
        return pk;
    }
    // chip.rs:58-70, RSAPubE::Var: the exponent is a witness too
    AssignedRSAPublicKey assign_public_key_var(RegionCtx& ctx, const UnassignedInteger& n, const UnassignedInteger& e) const {
        AssignedRSAPublicKey pk;
        BigIntChip chip = bigint_chip();
        pk.n = chip.assign_integer(ctx, n);
        pk.e_var = chip.assign_integer(ctx, e);
        return pk;
    }
    // chip.rs:80-88
    AssignedRSASignature assign_signature(RegionCtx& ctx, const UnassignedInteger& c) const {
        AssignedRSASignature s;
        s.c = bigint_chip().assign_integer(ctx, c);
        return s;
    }
    // chip.rs:99-114
    AssignedInteger modpow_public_key(RegionCtx& ctx, const AssignedInteger& x, const AssignedRSAPublicKey& pk) const {
        BigIntChip chip = bigint_chip();
        chip.assert_in_field(ctx, x, pk.n);
        if (pk.e_var.num_limbs()) return chip.pow_mod(ctx, x, pk.e_var, pk.n, exp_limb_bits);
        return chip.pow_mod_fixed_exp(ctx, x, pk.e_fix, pk.n);
    }
    // chip.rs:128-199
    AssignedValue verify_pkcs1v15_signature(RegionCtx& ctx, const AssignedRSAPublicKey& pk, const AssignedInteger& hashed_msg,
                                            const AssignedRSASignature& sig) const {
        MainGate mg;
        RangeChip rc;
        AssignedValue is_eq = mg.assign_constant(ctx, U256(1));
        AssignedInteger powed = modpow_public_key(ctx, sig.c, pk);
        const size_t hash_len = 4;
        for (size_t i = 0; i < hash_len; i++) {
            AssignedValue e = mg.is_equal(ctx, powed.limb(i), hashed_msg.limb(i));
            is_eq = mg.and_(ctx, is_eq, e);
        }
        AssignedValue prefix_64_1 = mg.assign_constant(ctx, U256(217300885422736416ull));
        AssignedValue prefix_64_2 = mg.assign_constant(ctx, U256(938447882527703397ull));
        AssignedValue e1 = mg.is_equal(ctx, powed.limb(hash_len), prefix_64_1);
        AssignedValue e2 = mg.is_equal(ctx, powed.limb(hash_len + 1), prefix_64_2);
        is_eq = mg.and_(ctx, is_eq, e1);
        is_eq = mg.and_(ctx, is_eq, e2);
        int32_t low_v = ctx.op1(OP_LOWBITS, powed.limb(hash_len + 2).vid, 32);
        int32_t high_v = ctx.op1(OP_SHR, powed.limb(hash_len + 2).vid, 32);
        AssignedValue remain_low = rc.assign(ctx, low_v, 4, 32);
        AssignedValue remain_high = rc.assign(ctx, high_v, 4, 32);
        AssignedValue u32_assign = mg.assign_constant(ctx, U256::pow2(32));
        AssignedValue remain_concat = mg.mul_add(ctx, remain_high, u32_assign, remain_low);
        mg.assert_equal(ctx, powed.limb(hash_len + 2), remain_concat);
        AssignedValue prefix_32 = mg.assign_constant(ctx, U256(3158320));
        AssignedValue e3 = mg.is_equal(ctx, remain_low, prefix_32);
        is_eq = mg.and_(ctx, is_eq, e3);
        AssignedValue ff_32 = mg.assign_constant(ctx, U256(4294967295ull));
        AssignedValue e4 = mg.is_equal(ctx, remain_high, ff_32);
        is_eq = mg.and_(ctx, is_eq, e4);
        AssignedValue ff_64 = mg.assign_constant(ctx, U256(18446744073709551615ull));
        for (size_t i = hash_len + 3; i < bits_len / LIMB_WIDTH - 1; i++) {
            AssignedValue e = mg.is_equal(ctx, powed.limb(i), ff_64);
            is_eq = mg.and_(ctx, is_eq, e);
        }
        AssignedValue last_em = mg.assign_constant(ctx, U256(562949953421311ull));
        AssignedValue e5 = mg.is_equal(ctx, powed.limb(bits_len / LIMB_WIDTH - 1), last_em);
        is_eq = mg.and_(ctx, is_eq, e5);
        return is_eq;
    }
};


// ---- the bench circuit's synthesize (benches/bench.rs:132-225, sha2 disabled) ---------------------
static constexpr uint32_t BLINDING_ROWS = 6;  // cs.blinding_factors() + 1 = 5 + 1 for this circuit

// RangeChip::configure with RSAChip::compute_range_lens: distinct non-zero bit lengths, ascending tags
inline void configure_range_tags(RegionCtx& rc, unsigned num_limbs) {
    std::vector<unsigned> comp, over, lens;
    RSAChip::compute_range_lens(num_limbs, comp, over);
    for (unsigned v : comp) if (v) lens.push_back(v);
    for (unsigned v : over) if (v) lens.push_back(v);
    std::sort(lens.begin(), lens.end());
    lens.erase(std::unique(lens.begin(), lens.end()), lens.end());
    for (size_t i = 0; i < lens.size(); i++) rc.tag_of_bits[lens[i]] = (int)i + 1;
}

// Records the whole circuit into `rc`; returns the is_valid cell.  Inputs are value nodes
// OP_INPUT 0..nl-1 = n limbs, nl..2nl-1 = signature limbs, 2nl..2nl+3 = hash limbs.
inline AssignedValue record_rsa_pkcs1v15(RegionCtx& rc, unsigned bits_len, const std::vector<uint8_t>& e_le) {
    const unsigned nl = bits_len / 64;
    configure_range_tags(rc, nl);
    RSAChip rsa_chip(bits_len, 5);
    BigIntChip bigint_chip = rsa_chip.bigint_chip();
    MainGate main_gate;
    UnassignedInteger sig_u, n_u, hash_u;
    for (unsigned i = 0; i < nl; i++) n_u.limbs.push_back(rc.input(i));
    for (unsigned i = 0; i < nl; i++) sig_u.limbs.push_back(rc.input(nl + i));
    for (unsigned i = 0; i < 4; i++) hash_u.limbs.push_back(rc.input(2 * nl + i));
    // region 1 (bench.rs:145-156): signature, then public key
    AssignedRSASignature sign = rsa_chip.assign_signature(rc, sig_u);
    AssignedRSAPublicKey public_key = rsa_chip.assign_public_key(rc, n_u, e_le);
    // region 2 (bench.rs:186-211): hashed message, verification
    AssignedInteger hashed = bigint_chip.assign_integer(rc, hash_u);
    AssignedValue is_valid = rsa_chip.verify_pkcs1v15_signature(rc, public_key, hashed, sign);
    // region 3 (bench.rs:213-221)
    main_gate.assert_one(rc, is_valid);
    return is_valid;
}

// pkcs1v15 circuit with RSAPubE::Var (src/chip.rs:372-390 test_rsa_signature_with_hash_circuit's shape): the exponent is an
// assigned one-limb integer of which `exp_limb_bits` bits are used.  Inputs: n limbs, signature limbs, hash limbs, then e.
// RSASignatureVerifier::verify_pkcs1v15_signature (src/lib.rs:183-248), the part the reference itself holds: the 32
// digest bytes that halo2-dynamic-sha256's chip hands over as assigned cells (`decompose_digest_to_bytes`, reversed:
// src/lib.rs:210-213) are composed into four 64-bit limbs - per limb `assign_constant(0)`, then for each of its 8 bytes
// `assign_constant(2^(8 j))` and `mul_add(coeff, byte, limb)` (src/lib.rs:222-236) - and the integer goes to
// RSAChip::verify_pkcs1v15_signature in the same region.  The SHA-256 compression itself lives in the external crate
// (unpinned, not vendored: DESIGN.md section 7); its output cells are stood in for by 32 `assign_value` rows in a region
// of their own, so everything from the byte cells on is the reference's layout.  No assert_one: the verifier returns
// is_valid to its caller (src/lib.rs:245).
inline AssignedValue record_rsa_verifier_from_digest(RegionCtx& rc, unsigned bits_len, const std::vector<uint8_t>& e_le) {
    const unsigned nl = bits_len / 64;
    configure_range_tags(rc, nl);
    RSAChip rsa_chip(bits_len, 5);
    MainGate main_gate;
    UnassignedInteger sig_u, n_u;
    for (unsigned i = 0; i < nl; i++) n_u.limbs.push_back(rc.input(i));
    for (unsigned i = 0; i < nl; i++) sig_u.limbs.push_back(rc.input(nl + i));
    AssignedRSASignature sign = rsa_chip.assign_signature(rc, sig_u);
    AssignedRSAPublicKey public_key = rsa_chip.assign_public_key(rc, n_u, e_le);
    // stand-in for the SHA chip's digest-byte cells: hashed_bytes[8 i + j] = byte j of limb i (least significant first)
    std::vector<AssignedValue> hashed_bytes;
    for (unsigned i = 0; i < 4; i++)
        for (unsigned j = 0; j < 8; j++) hashed_bytes.push_back(main_gate.assign_value(rc, rc.op1(OP_SUBLIMB, rc.input(2 * nl + i), 8 * j, 8)));
    // region "verify pkcs1v15 signature" (src/lib.rs:217-243)
    AssignedInteger hashed_msg;
    for (unsigned i = 0; i < 4; i++) {
        AssignedValue limb_val = main_gate.assign_constant(rc, U256(0));
        for (unsigned j = 0; j < 8; j++) {
            AssignedValue coeff = main_gate.assign_constant(rc, U256::pow2(8 * j));
            limb_val = main_gate.mul_add(rc, coeff, hashed_bytes[8 * i + j], limb_val);
        }
        hashed_msg.limbs.push_back(limb_val);
    }
    return rsa_chip.verify_pkcs1v15_signature(rc, public_key, hashed_msg, sign);
}

inline AssignedValue record_rsa_pkcs1v15_var(RegionCtx& rc, unsigned bits_len, unsigned exp_limb_bits) {
    const unsigned nl = bits_len / 64;
    configure_range_tags(rc, nl);
    RSAChip rsa_chip(bits_len, exp_limb_bits);
    BigIntChip bigint_chip = rsa_chip.bigint_chip();
    MainGate main_gate;
    UnassignedInteger sig_u, n_u, hash_u, e_u;
    for (unsigned i = 0; i < nl; i++) n_u.limbs.push_back(rc.input(i));
    for (unsigned i = 0; i < nl; i++) sig_u.limbs.push_back(rc.input(nl + i));
    for (unsigned i = 0; i < 4; i++) hash_u.limbs.push_back(rc.input(2 * nl + i));
    e_u.limbs.push_back(rc.input(2 * nl + 4));
    AssignedRSASignature sign = rsa_chip.assign_signature(rc, sig_u);
    AssignedRSAPublicKey public_key = rsa_chip.assign_public_key_var(rc, n_u, e_u);
    AssignedInteger hashed = bigint_chip.assign_integer(rc, hash_u);
    AssignedValue is_valid = rsa_chip.verify_pkcs1v15_signature(rc, public_key, hashed, sign);
    main_gate.assert_one(rc, is_valid);
    return is_valid;
}

// Single BigIntChip operations, as the reference's in-file test circuits drive them (src/big_integer/chip.rs:1470-2313).
// Inputs (value nodes): a = 0..nl-1, b = nl..2nl-1, n = 2nl..3nl-1, e = 3nl.  Returns the cell whose value reports the
// outcome (the last limb of the result); `out` receives the result limbs.
enum BigIntTestOp : uint32_t {
    BT_REFRESH = 6, BT_ADD_MOD = 7, BT_SUB_MOD = 8, BT_POW_MOD = 9,
    // the predicates (chip.rs:754-1006; test circuits chip.rs:2395-2795): the result integer is the one assigned bit
    BT_IS_ZERO = 10, BT_IS_EQUAL_FRESH = 11, BT_IS_LESS_THAN = 12, BT_IS_LESS_THAN_OR_EQUAL = 13, BT_IS_GREATER_THAN = 14,
    BT_IS_GREATER_THAN_OR_EQUAL = 15, BT_IS_IN_FIELD = 16,
    BT_SQUARE = 17, BT_SQUARE_MOD = 18   // chip.rs:431-437, 642-649
};
inline AssignedValue record_bigint_op(RegionCtx& rc, uint32_t op, unsigned bits_len, unsigned exp_limb_bits, AssignedInteger* out) {
    const unsigned nl = bits_len / 64;
    configure_range_tags(rc, nl);
    BigIntChip chip(64, bits_len);
    UnassignedInteger a_u, b_u, n_u, e_u;
    for (unsigned i = 0; i < nl; i++) a_u.limbs.push_back(rc.input(i));
    for (unsigned i = 0; i < nl; i++) b_u.limbs.push_back(rc.input(nl + i));
    for (unsigned i = 0; i < nl; i++) n_u.limbs.push_back(rc.input(2 * nl + i));
    e_u.limbs.push_back(rc.input(3 * nl));
    AssignedInteger a = chip.assign_integer(rc, a_u), r;
    if (op == BT_REFRESH) {            // chip.rs:1861-1899
        AssignedInteger b = chip.assign_integer(rc, b_u);
        AssignedInteger ab = chip.mul(rc, a, b), ba = chip.mul(rc, b, a);
        r = chip.refresh(rc, ab, nl, nl);
        AssignedInteger ba_r = chip.refresh(rc, ba, nl, nl);
        chip.assert_equal_fresh(rc, r, ba_r);
    } else if (op == BT_ADD_MOD || op == BT_SUB_MOD) {   // chip.rs:1948-2110
        AssignedInteger b = chip.assign_integer(rc, b_u);
        AssignedInteger n = chip.assign_integer(rc, n_u);
        r = op == BT_ADD_MOD ? chip.add_mod(rc, a, b, n) : chip.sub_mod(rc, a, b, n);
    } else if (op == BT_POW_MOD) {     // chip.rs:2229-2271 (the exponent assigned as a witness, as RSAPubE::Var does)
        AssignedInteger e = chip.assign_integer(rc, e_u);
        AssignedInteger n = chip.assign_integer(rc, n_u);
        r = chip.pow_mod(rc, a, e, n, exp_limb_bits);
    } else if (op >= BT_IS_ZERO && op <= BT_IS_IN_FIELD) {
        AssignedInteger b = chip.assign_integer(rc, b_u);
        AssignedValue bit;
        switch (op) {
            case BT_IS_ZERO: bit = chip.is_zero(rc, a); break;
            case BT_IS_EQUAL_FRESH: bit = chip.is_equal_fresh(rc, a, b); break;
            case BT_IS_LESS_THAN: bit = chip.is_less_than(rc, a, b); break;
            case BT_IS_LESS_THAN_OR_EQUAL: bit = chip.is_less_than_or_equal(rc, a, b); break;
            case BT_IS_GREATER_THAN: bit = chip.is_greater_than(rc, a, b); break;
            case BT_IS_GREATER_THAN_OR_EQUAL: bit = chip.is_greater_than_or_equal(rc, a, b); break;
            default: bit = chip.is_in_field(rc, a, b); break;
        }
        r.limbs.assign(1, bit);
    } else if (op == BT_SQUARE || op == BT_SQUARE_MOD) {
        AssignedInteger b = chip.assign_integer(rc, b_u);   // assigned and unused, to keep one input convention for every op
        if (op == BT_SQUARE) {
            r = chip.square(rc, a);
        } else {
            AssignedInteger n = chip.assign_integer(rc, n_u);
            r = chip.square_mod(rc, a, n);
        }
    } else {
        throw SynthError(-1, "record_bigint_op: unknown operation");
    }
    if (out) *out = r;
    return r.limbs.back();
}

}  // namespace circuit
}  // namespace b2r
