//! Expression lowering: lazy View tree -> position-space descriptor (`MdimExpr`).  UNVERIFIED — see rust/README.md.
//! A line-for-line restatement of multidimension_b200/lowering.py (+ `_refine` / `_unify_group` / `_sub_of` of
//! view.py), which IS tested: every index-remapping view is a rewrite of strides and offsets over named position axes.
use std::collections::HashMap;
use std::rc::Rc;

use crate::ffi::{self, MdimExpr, MdimNode};

/// Valid in the reference, but not lowerable to the device (MDIM_ERR_UNSUPPORTED).
#[derive(Debug)]
pub struct Unsupported(pub String);
pub type Lowering<T> = Result<T, Unsupported>;

/// One position-space axis: a leaf of a flattened index type (src/tuple.rs:60-176) with its run-time length.
#[derive(Copy, Clone, Debug, PartialEq, Eq, Hash)]
pub struct Axis { pub id: u32, pub length: u64 }

/// The position axes of ONE type leaf (more than one after `to_usize` merged several, src/view.rs:1029-1059).
pub type Group = Vec<Axis>;

/// DIAG pair: coord[a] == coord[b] + offset, or coord[a] == constant.
#[derive(Clone, Debug)]
pub enum Pair { Axes(Axis, Axis, i64), Const(Axis, u64) }

#[derive(Clone, Debug)]
pub struct Peers { pub ptr: Vec<*const core::ffi::c_void>, pub block: u64 }

/// One scalar-valued node.  `stride` maps Axis -> elements per step (absent = 0 = Broadcast, src/broadcast.rs:46-60).
#[derive(Clone, Debug)]
pub struct Node {
    pub kind: i32, pub dtype: i32, pub op: i32, pub src_dtype: i32,
    pub children: Vec<Rc<Node>>,
    pub data: *const core::ffi::c_void, pub offset: i64,
    pub stride: Vec<(Axis, i64)>,
    pub gstride: Vec<i64>, pub bound: Vec<u64>,
    pub pairs: Vec<Pair>,
    pub concat: Option<(Axis, u64)>,   // CONCAT: (axis, length of V along it)
    pub imm: u64,                      // raw little-endian bits of a value of `dtype`
    pub red_axes: Vec<Axis>,           // FOLD: reduction axes, iterated last-fastest
    pub peers: Option<Peers>,
}

impl Node {
    pub fn new(kind: i32, dtype: i32) -> Node {
        Node { kind, dtype, op: 0, src_dtype: 0, children: vec![], data: core::ptr::null(), offset: 0, stride: vec![], gstride: vec![],
               bound: vec![], pairs: vec![], concat: None, imm: 0, red_axes: vec![], peers: None }
    }
    pub fn stride_of(&self, a: Axis) -> i64 { self.stride.iter().find(|(x, _)| *x == a).map(|(_, s)| *s).unwrap_or(0) }
}

/// A VALUE is a scalar node or a tuple of values (the pairs of `zip`, src/ops.rs:25-29; compound `All`).
#[derive(Clone, Debug)]
pub enum Value { Scalar(Rc<Node>), Tuple(Vec<Value>) }

impl Value {
    pub fn map(&self, f: &mut dyn FnMut(&Rc<Node>) -> Lowering<Rc<Node>>) -> Lowering<Value> {
        Ok(match self {
            Value::Scalar(n) => Value::Scalar(f(n)?),
            Value::Tuple(vs) => Value::Tuple(vs.iter().map(|v| v.map(f)).collect::<Lowering<Vec<_>>>()?),
        })
    }
    pub fn leaves(&self) -> Vec<Rc<Node>> {
        match self { Value::Scalar(n) => vec![n.clone()], Value::Tuple(vs) => vs.iter().flat_map(|v| v.leaves()).collect() }
    }
}

/// What a view lowers to: per type-leaf groups of position axes (in `to_usize` order) and the element value tree.
pub struct Lowered { pub groups: Vec<Group>, pub value: Value }

pub struct Builder { next: u32 }
impl Builder {
    pub fn new() -> Builder { Builder { next: 0 } }
    pub fn axis(&mut self, length: u64) -> Axis { self.next += 1; Axis { id: self.next, length } }
}

/// Substitution of one old axis: coordinate = constant + sum(coef * new axis).
#[derive(Clone, Debug)]
pub struct Sub { pub constant: i64, pub terms: Vec<(Axis, i64)> }
impl Sub {
    pub fn pin(c: i64) -> Sub { Sub { constant: c, terms: vec![] } }
    pub fn rename(to: Axis) -> Sub { Sub { constant: 0, terms: vec![(to, 1)] } }
}

/// Row-major combination of the pieces an axis has been cut into (outermost first); unit pieces carry no term.
pub fn sub_of(pieces: &[Axis]) -> Sub {
    let (mut terms, mut acc) = (vec![], 1i64);
    for a in pieces.iter().rev() {
        if a.length != 1 { terms.push((*a, acc)); }
        acc *= a.length as i64;
    }
    Sub { constant: 0, terms }
}

/// Common refinement of two row-major factorisations of the same run of positions (view.py::_refine).
pub fn refine(a: &[u64], b: &[u64]) -> Lowering<(Vec<u64>, Vec<Vec<usize>>, Vec<Vec<usize>>)> {
    if a.contains(&0) || b.contains(&0) { return Err(Unsupported("re-splitting an empty axis".into())); }
    let (mut pieces, mut ma, mut mb) = (vec![], vec![vec![]; a.len()], vec![vec![]; b.len()]);
    let (mut i, mut j) = (0usize, 0usize);
    let (mut ra, mut rb): (Option<u64>, Option<u64>) = (None, None);
    loop {
        while ra.is_none() && i < a.len() { if a[i] == 1 { ma[i].push(pieces.len()); pieces.push(1); i += 1 } else { ra = Some(a[i]) } }
        while rb.is_none() && j < b.len() { if b[j] == 1 { mb[j].push(pieces.len()); pieces.push(1); j += 1 } else { rb = Some(b[j]) } }
        let (x, y) = match (ra, rb) { (Some(x), Some(y)) => (x, y), _ => break };
        let step = if x % y == 0 { y } else if y % x == 0 { x } else {
            return Err(Unsupported(format!("axis groups {:?} and {:?} have no common refinement: this remapping needs div/mod on the device", a, b)));
        };
        ma[i].push(pieces.len()); mb[j].push(pieces.len()); pieces.push(step);
        ra = if x / step == 1 { i += 1; None } else { Some(x / step) };
        rb = if y / step == 1 { j += 1; None } else { Some(y / step) };
    }
    if ra.is_some() || rb.is_some() { return Err(Unsupported(format!("axis groups {:?} and {:?} do not describe the same run of positions", a, b))); }
    Ok((pieces, ma, mb))
}

/// One leaf axis seen by both operands of a Zip / Concat (view.py::_unify_group).
pub fn unify_group(b: &mut Builder, ga: &Group, gb: &Group, tv: &mut HashMap<Axis, Sub>, tw: &mut HashMap<Axis, Sub>) -> Lowering<Group> {
    let (la, lb): (Vec<u64>, Vec<u64>) = (ga.iter().map(|a| a.length).collect(), gb.iter().map(|a| a.length).collect());
    if la == lb {
        for (x, y) in ga.iter().zip(gb) { tw.insert(*y, Sub::rename(*x)); }
        return Ok(ga.clone());
    }
    let (pieces, ma, mb) = refine(&la, &lb)?;
    let axes: Vec<Axis> = pieces.iter().map(|n| b.axis(*n)).collect();
    for (x, idx) in ga.iter().zip(&ma) { tv.insert(*x, sub_of(&idx.iter().map(|k| axes[*k]).collect::<Vec<_>>())); }
    for (y, idx) in gb.iter().zip(&mb) { tw.insert(*y, sub_of(&idx.iter().map(|k| axes[*k]).collect::<Vec<_>>())); }
    Ok(axes)
}

fn pred_side(x: Option<Axis>, c: i64, table: &HashMap<Axis, Sub>) -> Lowering<(Option<Axis>, i64)> {
    match x {
        None => Ok((None, c)),
        Some(a) => match table.get(&a) {
            None => Ok((Some(a), c)),
            Some(s) if s.terms.is_empty() => Ok((None, c + s.constant)),
            Some(s) if s.terms.len() == 1 && s.terms[0].1 == 1 => Ok((Some(s.terms[0].0), c + s.constant)),
            Some(_) => Err(Unsupported("a Diagonal whose axis has been split or merged needs device div/mod".into())),
        },
    }
}

/// Rewrite every reference to the axes in `table`.  Linear in strides / offsets (lowering.py::substitute):
/// stride'[n] += stride[a] * coef, offset' += stride[a] * constant.  Shared subtrees stay shared through `memo`.
pub fn substitute(node: &Rc<Node>, table: &HashMap<Axis, Sub>, memo: &mut HashMap<*const Node, Rc<Node>>) -> Lowering<Rc<Node>> {
    if table.is_empty() { return Ok(node.clone()); }
    if let Some(hit) = memo.get(&Rc::as_ptr(node)) { return Ok(hit.clone()); }
    let kids = node.children.iter().map(|c| substitute(c, table, memo)).collect::<Lowering<Vec<_>>>()?;
    let mut out = (**node).clone();
    out.children = kids.clone();
    match node.kind {
        ffi::LEAF | ffi::IOTA | ffi::GATHER => {
            let mut stride: Vec<(Axis, i64)> = vec![];
            let mut add = |stride: &mut Vec<(Axis, i64)>, a: Axis, s: i64| match stride.iter_mut().find(|(x, _)| *x == a) { Some(e) => e.1 += s, None => stride.push((a, s)) };
            for (a, s) in &node.stride {
                match table.get(a) {
                    None => add(&mut stride, *a, *s),
                    Some(sub) => { out.offset += s * sub.constant; for (n, coef) in &sub.terms { add(&mut stride, *n, s * coef); } }
                }
            }
            stride.retain(|(_, s)| *s != 0);
            out.stride = stride;
        }
        ffi::DIAG => {
            let (mut pairs, mut dead) = (vec![], false);
            for pr in &node.pairs {
                let (a, b, off) = match pr { Pair::Axes(a, b, off) => (Some(*a), Some(*b), *off), Pair::Const(a, k) => (Some(*a), None, *k as i64) };
                let (a2, ca) = pred_side(a, 0, table)?;
                let (b2, cb) = pred_side(b, off, table)?;
                match (a2, b2) {
                    (None, None) => { if ca != cb { dead = true; } }
                    (Some(x), None) => { let k = cb - ca; if k < 0 { dead = true } else { pairs.push(Pair::Const(x, k as u64)) } }
                    (None, Some(y)) => { let k = ca - cb; if k < 0 { dead = true } else { pairs.push(Pair::Const(y, k as u64)) } }
                    (Some(x), Some(y)) => pairs.push(Pair::Axes(x, y, cb - ca)),
                }
            }
            if dead { let mut c = Node::new(ffi::CONST, node.dtype); c.imm = node.imm; out = c; }   // statically off the diagonal
            else if pairs.is_empty() { out = (*kids[0]).clone(); }
            else { out.pairs = pairs; }
        }
        ffi::CONCAT => {
            let (axis, thr) = node.concat.unwrap();
            if let Some(sub) = table.get(&axis) {
                if sub.terms.is_empty() { out = (*kids[if (sub.constant as u64) < thr { 0 } else { 1 }]).clone(); }   // pinned: one side survives
                else if sub.terms.len() == 1 && sub.terms[0].1 == 1 { out.concat = Some((sub.terms[0].0, (thr as i64 - sub.constant).max(0) as u64)); }
                else { return Err(Unsupported("a Concat whose axis has been split needs device div/mod".into())); }
            }
        }
        ffi::FOLD => { if node.red_axes.iter().any(|a| table.contains_key(a)) { return Err(Unsupported("substitution of a reduction axis".into())); } }
        _ => {}
    }
    let out = Rc::new(out);
    memo.insert(Rc::as_ptr(node), out.clone());
    Ok(out)
}

/// The emitted descriptor and the storage that must outlive the call.
pub struct Emitted { pub nodes: Vec<MdimNode>, pub expr: MdimExpr, pub out_dtype: i32, pub out_len: u64 }

/// Number the root's axes in `to_usize` order and write the tree in post-order (lowering.py::emit).
pub fn emit(root: &Rc<Node>, axes: &[Axis]) -> Lowering<Emitted> {
    fn visit(n: &Rc<Node>, order: &mut Vec<Rc<Node>>, red: &mut Vec<Axis>) -> Lowering<()> {
        for c in &n.children { visit(c, order, red)?; }
        if n.kind == ffi::FOLD {
            if !red.is_empty() { return Err(Unsupported("more than one fold in one expression".into())); }
            red.extend(n.red_axes.iter().copied());
        }
        order.push(n.clone());
        Ok(())
    }
    let (mut order, mut red) = (vec![], vec![]);
    visit(root, &mut order, &mut red)?;
    if order.len() > ffi::MDIM_MAX_NODES { return Err(Unsupported(format!("expression has {} nodes (> {})", order.len(), ffi::MDIM_MAX_NODES))); }
    let all: Vec<Axis> = axes.iter().chain(red.iter()).copied().collect();
    if all.len() > ffi::MDIM_MAX_RANK { return Err(Unsupported(format!("expression has {} position axes (> {})", all.len(), ffi::MDIM_MAX_RANK))); }
    let pos = |a: &Axis| all.iter().position(|x| x == a).ok_or_else(|| Unsupported(format!("internal: node strides over an axis that is not iterated ({:?})", a)));
    let mut nodes = Vec::with_capacity(order.len());
    for n in &order {
        let mut d = MdimNode::zeroed();
        d.kind = n.kind; d.dtype = n.dtype; d.op = n.op; d.src_dtype = n.src_dtype;
        if matches!(n.kind, ffi::LEAF | ffi::IOTA | ffi::GATHER) {
            d.offset = n.offset;
            for (a, s) in &n.stride { d.stride[pos(a)?] = *s; }
        }
        if matches!(n.kind, ffi::LEAF | ffi::GATHER) {
            match &n.peers {
                Some(p) => { d.n_peers = p.ptr.len() as i32; for (k, q) in p.ptr.iter().enumerate() { d.peer[k] = *q; } d.peer_block = p.block; }
                None => d.data = n.data,
            }
        }
        if n.kind == ffi::GATHER {
            d.n_comp = n.children.len() as i32;
            for c in 0..n.children.len() { d.gstride[c] = n.gstride[c]; d.bound[c] = n.bound[c]; }
        }
        if n.kind == ffi::DIAG {
            d.n_comp = n.pairs.len() as i32;
            for (p, pr) in n.pairs.iter().enumerate() {
                match pr {
                    Pair::Axes(a, b, off) => { d.axis_a[p] = pos(a)? as i32; d.axis_b[p] = pos(b)? as i32; d.axis_c[p] = *off as u64; }
                    Pair::Const(a, k) => { d.axis_a[p] = pos(a)? as i32; d.axis_b[p] = -1; d.axis_c[p] = *k; }
                }
            }
            d.imm.u64_ = n.imm;
        }
        if let (ffi::CONCAT, Some((a, thr))) = (n.kind, n.concat) { d.axis_a[0] = pos(&a)? as i32; d.axis_c[0] = thr; }
        if matches!(n.kind, ffi::CONST | ffi::FOLD) { d.imm.u64_ = n.imm; }
        nodes.push(d);
    }
    let mut length = [0u64; ffi::MDIM_MAX_RANK];
    for (i, a) in all.iter().enumerate() { length[i] = a.length; }
    let expr = MdimExpr { abi_version: ffi::MDIM_ABI_VERSION, rank: axes.len() as i32, red_rank: red.len() as i32, n_nodes: nodes.len() as i32,
                          length, nodes: nodes.as_ptr() };
    Ok(Emitted { nodes, expr, out_dtype: root.dtype, out_len: axes.iter().map(|a| a.length).product() })
}

/// The position-space form of a view.  Implemented next to every node struct (rust/patches/view_rs.rs), because the
/// node structs' fields are private tuple fields of src/view.rs.
pub trait Lower: crate::View {
    fn lower(&self, b: &mut Builder) -> Lowering<Lowered>;
}
