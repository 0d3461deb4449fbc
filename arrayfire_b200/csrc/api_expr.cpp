// api_expr.cpp -- SURVEY.md section 8(f)-4: lazily evaluated element-wise expressions that a whole-array reduction fuses
// into its load path (kernels/expr.cu).
//
// What it mirrors in the reference: af_add / af_mul / af_lt / af_sqrt ... return JIT nodes instead of evaluated arrays
// (src/api/c/binary.cpp:40-120, unary.cpp, src/backend/common/jit/{Node,BinaryNode,UnaryNode,ScalarNode}.hpp) and a
// consumer evaluates the tree in one kernel; af_constant is a ScalarNode (no memory, src/api/c/data.cpp:45-90).  The
// pi benchmark (examples/benchmarks/pi.cpp:28-31) is the shape this serves:  sum<float>(sqrt(x*x + y*y) < 1).
// Scope: f32 / f64 / b8 operands of identical dims; arithmetic (+ - * /), comparisons, logical and / or / not, sqrt, abs,
// min / max / pow; s32 / u32 / s16 / u16 / s8 / u8 operands in comparisons and logic (exact in the double interpreter).  Anything else returns AF_ERR_NOT_SUPPORTED (never a wrong result, never a CPU fallback).
// Fused consumers: af_sum_all, af_count_all, af_min_all, af_max_all, af_any_true_all, af_all_true_all (api_algo.cpp);
// every other entry point sees a normal array because from_handle() materialises the expression first.
#include <cmath>
#include <vector>

#include "common.hpp"

using namespace afb;

namespace afb {
Array *make_dense_copy(const Array &a);

struct ExprNode {
    int op = EX_LEAF;
    af_dtype type = f32;                     // dtype of this node's value
    std::shared_ptr<ExprNode> kid[2];
    std::shared_ptr<Buffer> buf;             // leaf
    size_t offset_bytes = 0;
    double cst = 0.0;                        // constant
    bool via_f32 = false;                    // integer leaf whose consumer works in f32: value rounded through float
    int nops = 1;                            // instructions of the flattened subtree
};
}  // namespace afb

namespace {

struct OpLock {
    std::lock_guard<std::mutex> g;
    explicit OpLock(int device) : g(op_mutex(device)) {}
};
struct Holder {
    Array *a = nullptr;
    ~Holder() { delete a; }
    Array *release() { Array *r = a; a = nullptr; return r; }
};

bool expr_type_ok(af_dtype t) { return t == f32 || t == f64 || t == b8; }
// integer operands: comparisons / logic only (the value is exact in the double interpreter); no integer-valued node exists
bool expr_int_ok(af_dtype t) { return t == s32 || t == u32 || t == s16 || t == u16 || t == s8 || t == u8; }
bool is_signed_int(af_dtype t) { return t == s32 || t == s16 || t == s8; }
int leaf_code(af_dtype t) {
    switch (t) {
        case f32: return 0;
        case f64: return 1;
        case s32: return 3;
        case u32: return 4;
        case s16: return 5;
        case u16: return 6;
        case u8: return 7;
        default: return 2;   // b8, s8: one signed byte
    }
}
// the reference's implicit() (src/api/c/implicit.cpp:20-45) restricted to the operand types of this interpreter
af_dtype implicit_of(af_dtype l, af_dtype r) {
    if (l == r) return l;
    if (l == f64 || r == f64) return f64;
    if (l == f32 || r == f32) return f32;
    const af_dtype order[] = {u32, s32, u16, s16, u8, s8};
    for (af_dtype t : order)
        if (l == t || r == t) return t;
    return b8;
}

// the node an operand contributes: its own expression, or a leaf over its (dense) buffer
std::shared_ptr<ExprNode> node_of(Array &a, Holder &keep) {
    if (a.expr) return a.expr;
    const Array *src = &a;
    if (!a.is_linear()) {
        keep.a = make_dense_copy(a);
        src = keep.a;
    }
    auto n = std::make_shared<ExprNode>();
    n->op = EX_LEAF;
    n->type = a.info.type;
    n->buf = src->buf;
    n->offset_bytes = (size_t)src->info.offset * dtype_size(src->info.type);
    return n;
}

// post-order flattening; returns the stack depth the subtree needs (entries besides the register-resident top)
int flatten(const ExprNode &n, ExprProg &p, int &nleaf, int &ncst, bool &dbl) {
    if (p.n >= kExprMaxOps) fail(AF_ERR_INTERNAL, "expression program too long");
    if (n.op == EX_LEAF) {
        const void *ptr = (const char *)n.buf->ptr + n.offset_bytes;
        int idx = -1;
        const int lt = leaf_code(n.type) | (n.via_f32 ? 16 : 0);
        for (int i = 0; i < nleaf; i++)
            if (p.leaf[i] == ptr && p.leaf_type[i] == lt) idx = i;
        if (idx < 0) {
            if (nleaf >= kExprMaxLeaves) fail(AF_ERR_INTERNAL, "expression has too many operands");
            idx = nleaf++;
            p.leaf[idx] = ptr;
            p.leaf_type[idx] = lt;
            if (reinterpret_cast<uintptr_t>(ptr) & 15) p.aligned = 0;
        }
        if (n.type == f64 || lt >= 3) dbl = true;
        p.op[p.n] = EX_LEAF;
        p.arg[p.n++] = (unsigned char)idx;
        return 0;
    }
    if (n.op == EX_CONST) {
        int idx = -1;
        for (int i = 0; i < ncst; i++)
            if (p.cst[i] == n.cst) idx = i;
        if (idx < 0) {
            if (ncst >= kExprMaxConsts) fail(AF_ERR_INTERNAL, "expression has too many constants");
            idx = ncst++;
            p.cst[idx] = n.cst;
        }
        if (n.type == f64 || expr_int_ok(n.type)) dbl = true;
        p.op[p.n] = EX_CONST;
        p.arg[p.n++] = (unsigned char)idx;
        return 0;
    }
    auto const_index = [&](const ExprNode &c) {
        int idx = -1;
        for (int i = 0; i < ncst; i++)
            if (p.cst[i] == c.cst) idx = i;
        if (idx < 0) {
            if (ncst >= kExprMaxConsts) fail(AF_ERR_INTERNAL, "expression has too many constants");
            idx = ncst++;
            p.cst[idx] = c.cst;
        }
        if (c.type == f64 || expr_int_ok(c.type)) dbl = true;
        return idx;
    };
    int depth, flags = 0, arg = 0;
    if (n.op < EX_FIRST_UNARY && n.kid[1]->op == EX_CONST && n.kid[0]->op != EX_CONST) {
        depth = flatten(*n.kid[0], p, nleaf, ncst, dbl);      // x op c: the scalar rides in the instruction
        flags = EX_RCONST;
        arg = const_index(*n.kid[1]);
    } else if (n.op < EX_FIRST_UNARY && n.kid[0]->op == EX_CONST && n.kid[1]->op != EX_CONST) {
        depth = flatten(*n.kid[1], p, nleaf, ncst, dbl);      // c op x
        flags = EX_LCONST;
        arg = const_index(*n.kid[0]);
    } else {
        depth = flatten(*n.kid[0], p, nleaf, ncst, dbl);
        if (n.op < EX_FIRST_UNARY) {
            const int right = flatten(*n.kid[1], p, nleaf, ncst, dbl) + 1;   // the left value waits in shared memory meanwhile
            if (right > depth) depth = right;
        }
    }
    if (n.type == f64) dbl = true;
    if (p.n >= kExprMaxOps) fail(AF_ERR_INTERNAL, "expression program too long");
    p.op[p.n] = (unsigned char)(n.op | flags);
    p.arg[p.n++] = (unsigned char)arg;
    return depth;
}

ExprProg program_of(const ExprNode &root, bool *dbl) {
    ExprProg p;
    memset(&p, 0, sizeof(p));
    p.aligned = 1;
    int nleaf = 0, ncst = 0;
    *dbl = false;
    const int depth = flatten(root, p, nleaf, ncst, *dbl);
    p.nleaf = nleaf;
    p.depth = depth;
    if (depth > 8) fail(AF_ERR_INTERNAL, "expression nests too deeply");   // builders evaluate sub-trees before this happens
    return p;
}

int depth_of(const ExprNode &n) {
    if (n.op == EX_LEAF || n.op == EX_CONST) return 0;
    int d = depth_of(*n.kid[0]);
    if (n.op < EX_FIRST_UNARY) {
        const int r = depth_of(*n.kid[1]) + 1;
        if (r > d) d = r;
    }
    return d;
}
int leaves_of(const ExprNode &n) {
    if (n.op == EX_LEAF || n.op == EX_CONST) return 1;
    return leaves_of(*n.kid[0]) + (n.op < EX_FIRST_UNARY ? leaves_of(*n.kid[1]) : 0);
}

Array *lazy_array(const Array &like, af_dtype type, std::shared_ptr<ExprNode> node) {
    Array *a = new Array();
    a->info = like.info;
    a->info.type = type;
    a->info.offset = 0;
    dense_strides(a->info.dims, a->info.strides);
    a->expr = std::move(node);
    return a;
}

// keeps programs within the interpreter's limits: a sub-tree that has grown too large is evaluated into an array first
std::shared_ptr<ExprNode> bounded(std::shared_ptr<ExprNode> n, const Array &like) {
    if (n->nops <= kExprMaxOps / 2 - 2 && depth_of(*n) <= 6 && leaves_of(*n) <= 6) return n;
    Holder tmp;
    tmp.a = lazy_array(like, n->type, n);
    materialize(*tmp.a);
    auto leaf = std::make_shared<ExprNode>();
    leaf->op = EX_LEAF;
    leaf->type = n->type;
    leaf->buf = tmp.a->buf;
    return leaf;
}

enum Kind { K_ARITH, K_COMPARE, K_LOGIC };

af_err binary(af_array *out, const af_array lhs_h, const af_array rhs_h, int op, Kind kind, const char *name) {
    AFB_API_BEGIN
    AFB_ARG(0, out != nullptr);
    Array *l = peek_handle(lhs_h), *r = peek_handle(rhs_h);
    for (int i = 0; i < 4; i++)
        if (l->info.dims[i] != r->info.dims[i])
            fail(AF_ERR_NOT_SUPPORTED, "%s: operands must have identical dims (batch / broadcast is not supported by libafb200)", name);
    const af_dtype tl = l->info.type, tr = r->info.type;
    const bool ints = expr_int_ok(tl) || expr_int_ok(tr);
    if (!(expr_type_ok(tl) || expr_int_ok(tl)) || !(expr_type_ok(tr) || expr_int_ok(tr)))
        fail(AF_ERR_NOT_SUPPORTED, "%s: libafb200 evaluates element-wise expressions over f32 / f64 / b8 operands (and 8- to 32-bit integers in comparisons) only", name);
    const af_dtype imp = implicit_of(tl, tr);
    if (ints) {
        if (kind == K_ARITH)
            fail(AF_ERR_NOT_SUPPORTED, "%s: integer operands are supported in comparisons and logical operations only (af_cast them for arithmetic)", name);
        // the reference compares in implicit(l, r): an unsigned result type would wrap a signed operand's negatives
        if ((imp == u32 || imp == u16 || imp == u8) && (is_signed_int(tl) || is_signed_int(tr)))
            fail(AF_ERR_NOT_SUPPORTED, "%s: mixed signed / unsigned integer operands are not supported", name);
    }
    check_same_device(*l);
    check_same_device(*r);
    af_dtype to;
    if (kind == K_ARITH) {
        if (l->info.type == b8 && r->info.type == b8) fail(AF_ERR_NOT_SUPPORTED, "%s: arithmetic on two b8 operands is not supported", name);
        to = (l->info.type == f64 || r->info.type == f64) ? f64 : f32;     // implicit(): the wider floating type
    } else {
        to = b8;
    }
    Holder kl, kr;
    auto n = std::make_shared<ExprNode>();
    n->type = to;
    // an s32 / u32 operand meeting an f32 one is cast to float by the reference before the comparison
    auto adapt = [&](std::shared_ptr<ExprNode> k, af_dtype own) {
        if (!(imp == f32 && (own == s32 || own == u32))) return k;
        auto c = std::make_shared<ExprNode>(*k);
        if (c->op == EX_LEAF) c->via_f32 = true;
        else if (c->op == EX_CONST) c->cst = (double)(float)c->cst;
        return c;
    };
    n->kid[0] = adapt(bounded(node_of(*l, kl), *l), tl);
    // x * x of the SAME operand (same handle, or the same buffer region): one load and a square
    const bool same = (l == r) || (l->expr && l->expr == r->expr) ||
                      (!l->expr && !r->expr && l->buf == r->buf && l->info.offset == r->info.offset && l->is_linear() && r->is_linear());
    if (op == EX_MUL && same && l->info.type == r->info.type) {
        n->op = EX_SQR;
        n->nops = n->kid[0]->nops + 1;
    } else {
        n->op = op;
        n->kid[1] = adapt(bounded(node_of(*r, kr), *r), tr);
        n->nops = n->kid[0]->nops + n->kid[1]->nops + 1;
    }
    *out = lazy_array(*l, to, n);
    AFB_API_END
}

af_err unary(af_array *out, const af_array in_h, int op, bool to_bool, const char *name) {
    AFB_API_BEGIN
    AFB_ARG(0, out != nullptr);
    Array *in = peek_handle(in_h);
    if (!(expr_type_ok(in->info.type) || (to_bool && expr_int_ok(in->info.type))) || (!to_bool && in->info.type == b8))
        fail(AF_ERR_NOT_SUPPORTED, "%s: libafb200 evaluates this function over f32 / f64 operands only", name);
    check_same_device(*in);
    Holder k;
    auto n = std::make_shared<ExprNode>();
    n->op = op;
    n->type = to_bool ? b8 : in->info.type;
    n->kid[0] = bounded(node_of(*in, k), *in);
    n->nops = n->kid[0]->nops + 1;
    *out = lazy_array(*in, n->type, n);
    AFB_API_END
}

}  // namespace

namespace afb {

void materialize(Array &a) {
    if (!a.expr) return;
    std::shared_ptr<ExprNode> root = a.expr;
    bool dbl = false;
    const ExprProg p = program_of(*root, &dbl);
    const dim_t n = a.elements();
    const af_dtype t = a.info.type;
    std::shared_ptr<Buffer> buf = pool_alloc((size_t)(n > 0 ? n : 1) * dtype_size(t));
    if (n > 0) {
        OpLock lk(a.device());
        LaunchCtx c = ctx_for_active_device();
        if (expr_int_ok(t)) launch_fill(c, t, buf->ptr, n, root->cst, 0.0);   // the only integer-valued node is a constant
        else launch_expr_eval(c, p, dbl, buf->ptr, t == f32 ? 0 : (t == f64 ? 1 : 2), n);
    }
    a.buf = buf;
    a.info.offset = 0;
    dense_strides(a.info.dims, a.info.strides);
    a.expr.reset();
}

// a lazy constant (af_constant of f32 / f64 / b8 and the 8- to 32-bit integers): no memory until somebody needs the values
Array *lazy_constant(af_dtype type, const dim_t dims[4], double val) {
    Array *a = new_array(type, (const dim_t[4]){0, 1, 1, 1});   // a handle with this device's identity, no payload
    for (int i = 0; i < 4; i++) a->info.dims[i] = dims[i];
    dense_strides(a->info.dims, a->info.strides);
    a->buf.reset();
    auto n = std::make_shared<ExprNode>();
    n->op = EX_CONST;
    n->type = type;
    switch (type) {      // the value the stored type holds (launch_fill's conversions)
        case b8: n->cst = val != 0.0 ? 1.0 : 0.0; break;
        case f32: n->cst = (double)(float)val; break;
        case s32: n->cst = (double)(int)val; break;
        case u32: n->cst = (double)(unsigned int)val; break;
        case s16: n->cst = (double)(short)val; break;
        case u16: n->cst = (double)(unsigned short)val; break;
        case s8: n->cst = (double)(signed char)val; break;
        case u8: n->cst = (double)(unsigned char)val; break;
        default: n->cst = val; break;
    }
    a->expr = n;
    return a;
}

// whole-array reduction of a lazy array, fused: mode EXR_*; the scalar lands in ret.host (f32 / f64 / u32 / b8)
bool expr_reduce_all(const Array &a, int mode, af_dtype *result_type, HostRet *ret_out) {
    if (!a.expr) return false;
    bool dbl = false;
    const ExprProg p = program_of(*a.expr, &dbl);
    const dim_t n = a.elements();
    OpLock lk(a.device());
    LaunchCtx c = ctx_for_active_device();
    HostRet ret = c.scratch->next_host_ret();
    launch_expr_reduce(c, p, dbl, mode, n, ret.dev, &ret);
    *result_type = (mode == EXR_COUNT) ? u32 : ((mode == EXR_ANY || mode == EXR_ALL) ? b8 : (dbl ? f64 : f32));
    *ret_out = ret;
    return true;
}

}  // namespace afb

extern "C" {

af_err af_add(af_array *o, const af_array l, const af_array r, const bool) { return binary(o, l, r, EX_ADD, K_ARITH, "af_add"); }
af_err af_sub(af_array *o, const af_array l, const af_array r, const bool) { return binary(o, l, r, EX_SUB, K_ARITH, "af_sub"); }
af_err af_mul(af_array *o, const af_array l, const af_array r, const bool) { return binary(o, l, r, EX_MUL, K_ARITH, "af_mul"); }
af_err af_div(af_array *o, const af_array l, const af_array r, const bool) { return binary(o, l, r, EX_DIV, K_ARITH, "af_div"); }
af_err af_minof(af_array *o, const af_array l, const af_array r, const bool) { return binary(o, l, r, EX_MIN, K_ARITH, "af_minof"); }
af_err af_maxof(af_array *o, const af_array l, const af_array r, const bool) { return binary(o, l, r, EX_MAX, K_ARITH, "af_maxof"); }
af_err af_pow(af_array *o, const af_array l, const af_array r, const bool) { return binary(o, l, r, EX_POW, K_ARITH, "af_pow"); }
af_err af_lt(af_array *o, const af_array l, const af_array r, const bool) { return binary(o, l, r, EX_LT, K_COMPARE, "af_lt"); }
af_err af_gt(af_array *o, const af_array l, const af_array r, const bool) { return binary(o, l, r, EX_GT, K_COMPARE, "af_gt"); }
af_err af_le(af_array *o, const af_array l, const af_array r, const bool) { return binary(o, l, r, EX_LE, K_COMPARE, "af_le"); }
af_err af_ge(af_array *o, const af_array l, const af_array r, const bool) { return binary(o, l, r, EX_GE, K_COMPARE, "af_ge"); }
af_err af_eq(af_array *o, const af_array l, const af_array r, const bool) { return binary(o, l, r, EX_EQ, K_COMPARE, "af_eq"); }
af_err af_neq(af_array *o, const af_array l, const af_array r, const bool) { return binary(o, l, r, EX_NEQ, K_COMPARE, "af_neq"); }
af_err af_and(af_array *o, const af_array l, const af_array r, const bool) { return binary(o, l, r, EX_AND, K_LOGIC, "af_and"); }
af_err af_or(af_array *o, const af_array l, const af_array r, const bool) { return binary(o, l, r, EX_OR, K_LOGIC, "af_or"); }
af_err af_sqrt(af_array *o, const af_array in) { return unary(o, in, EX_SQRT, false, "af_sqrt"); }
af_err af_abs(af_array *o, const af_array in) { return unary(o, in, EX_ABS, false, "af_abs"); }
af_err af_not(af_array *o, const af_array in) { return unary(o, in, EX_NOT, true, "af_not"); }

}  // extern "C"
