// ---- append to src/view.rs -----------------------------------------------------------------------------------------------
// UNVERIFIED — see rust/README.md.  One `impl Lower` per node struct, mirroring multidimension_b200/view.py::<Node>._lower
// rule by rule (which the differential test pins against the index-tuple-level model on GPU).  The impls live HERE because
// the node structs' fields are private tuple fields (`Transpose<V,I,X,Y,J>(V, PhantomData<..>)`, src/view.rs:1266, …).
//
// (1) The one change to the sink protocol (src/view.rs:20-38, 146-150): stable Rust has no specialisation, so `NewView` gets a
//     provided hook that defaults to today's element-at-a-time path, and `collect` calls it.
//
//   pub trait NewView: View {
//       type Buffer: Push<Self::T>;
//       fn new_view(size: <Self::I as Index>::Size, callback: impl FnOnce(&mut Self::Buffer)) -> Self;
//       /// NEW.  Default = src/view.rs:146-150 as it is today.  DeviceArray overrides it (rust/src/device_array.rs).
//       fn from_view<V: View<I = Self::I, T = Self::T> + MaybeLower>(v: &V) -> Self {
//           Self::new_view(v.size(), |buffer| v.each(|t| buffer.push(t)))
//       }
//   }
//   fn collect<A: NewView<I = Self::I, T = Self::T>>(&self) -> A { A::from_view(self) }          // View::collect

use std::collections::HashMap;
use std::rc::Rc;
use crate::device_ops::{binary_node, cast_node, const_node, unary_node, DeviceBinary, DeviceElem, DeviceUnary, Fold, Sym};
use crate::ffi;
use crate::lower::{self, sub_of, substitute, unify_group, Axis, Builder, Group, Lower, Lowered, Lowering, Node, Pair, Sub, Unsupported, Value};

fn flat(groups: &[Group]) -> Vec<Axis> { groups.iter().flatten().copied().collect() }

/// Partition the per-leaf groups of a view indexed by something isomorphic to a tuple of types with these leaf counts.
fn split_groups(groups: Vec<Group>, counts: &[usize]) -> Vec<Vec<Group>> {
    let mut it = groups.into_iter();
    counts.iter().map(|n| (0..*n).map(|_| it.next().expect("isomorphic index types have the same leaves")).collect()).collect()
}

fn subst_value(v: &Value, table: &HashMap<Axis, Sub>) -> Lowering<Value> {
    let mut memo = HashMap::new();
    v.map(&mut |n| substitute(n, table, &mut memo))
}

/// Pin the axes of `groups` at the positions of an index value (Row / Column); a leaf whose group has been re-split
/// (to_usize / from_usize) is pinned through its linear position, digit by digit (view.py::_pin).
fn pin<I: Index>(groups: &[Group], index: I, size: I::Size) -> HashMap<Axis, Sub> {
    let (mut pos, mut lens) = (vec![], vec![]);
    index.leaf_positions(size, &mut pos);
    I::leaf_lengths(size, &mut lens);
    let mut table = HashMap::new();
    for ((g, p), l) in groups.iter().zip(&pos).zip(&lens) {
        let same = g.len() == l.len() && g.iter().zip(l).all(|(a, n)| a.length == *n);
        let digits: Vec<u64> = if same { p.clone() } else {
            let mut k = p.iter().zip(l).fold(0u64, |k, (p, n)| k * n + p);
            let mut d: Vec<u64> = g.iter().rev().map(|a| { let r = k % a.length.max(1); k /= a.length.max(1); r }).collect();
            d.reverse();
            d
        };
        for (a, d) in g.iter().zip(digits) { table.insert(*a, Sub::pin(d as i64)); }
    }
    table
}

// ---- Scalar (src/view.rs:1399-1408): CONST -------------------------------------------------------------------------------------
impl<T: DeviceElem> Lower for Scalar<T> {
    fn lower(&self, _b: &mut Builder) -> Lowering<Lowered> { Ok(Lowered { groups: vec![], value: Value::Scalar(const_node(T::DTYPE, self.0.bits())) }) }
}

// ---- All (src/index.rs:177-186): IOTA per leaf; the value is shaped like I ---------------------------------------------------
fn iota_of(groups: &[Group]) -> Vec<Rc<Node>> {
    groups.iter().map(|g| {
        let mut n = Node::new(ffi::IOTA, ffi::U64);
        let mut acc = 1i64;
        for a in g.iter().rev() { n.stride.push((*a, acc)); acc *= a.length as i64; }   // a re-split leaf: its row-major position
        Rc::new(n)
    }).collect()
}
impl<I: Index> Lower for All<I> {   // element type I: scalar leaves only (usize / bool); compound I collects as a structure of arrays
    fn lower(&self, b: &mut Builder) -> Lowering<Lowered> {
        let mut lens = vec![];
        I::leaf_lengths(self.0, &mut lens);
        let groups: Vec<Group> = lens.iter().map(|g| g.iter().map(|n| b.axis(*n)).collect()).collect();
        let nodes = iota_of(&groups);
        let value = if nodes.len() == 1 { Value::Scalar(nodes[0].clone()) } else { Value::Tuple(nodes.into_iter().map(Value::Scalar).collect()) };
        Ok(Lowered { groups, value })
    }
}

// ---- Enumerate (src/view.rs:829-838): (index, value) ------------------------------------------------------------------------------
impl<V: Lower> Lower for Enumerate<V> {
    fn lower(&self, b: &mut Builder) -> Lowering<Lowered> {
        let l = self.0.lower(b)?;
        let idx = Value::Tuple(iota_of(&l.groups).into_iter().map(Value::Scalar).collect());
        Ok(Lowered { value: Value::Tuple(vec![idx, l.value]), groups: l.groups })
    }
}

// ---- Diagonal (src/view.rs:846-857): twin axes + DIAG predicate; the inner view is NOT evaluated off the diagonal -------
impl<V: Lower> Lower for Diagonal<V> where V::T: DeviceElem {
    fn lower(&self, b: &mut Builder) -> Lowering<Lowered> {
        let l = self.0.lower(b)?;
        let twin: Vec<Group> = l.groups.iter().map(|g| g.iter().map(|a| b.axis(a.length)).collect()).collect();
        let pairs: Vec<Pair> = flat(&l.groups).into_iter().zip(flat(&twin)).map(|(x, y)| Pair::Axes(x, y, 0)).collect();
        let zero = self.1.bits();
        let value = l.value.map(&mut |n| {
            if pairs.is_empty() { return Ok(n.clone()); }
            let mut d = Node::new(ffi::DIAG, n.dtype);
            d.children = vec![n.clone()]; d.pairs = pairs.clone(); d.imm = zero;
            Ok(Rc::new(d))
        })?;
        let mut groups = l.groups;
        groups.extend(twin);
        Ok(Lowered { groups, value })
    }
}

// ---- Map (src/view.rs:880-889): the closed unary vocabulary, `Fold<B>` over rows(), or a closure traced over Sym<T> ----------
pub struct Unary<V, U>(V, PhantomData<U>);                    // returned by the new `View::unary::<U>()`, beside `binary` (src/view.rs:507-512)
impl<V: Lower, U: DeviceUnary> Lower for Unary<V, U> where V::T: DeviceElem {
    fn lower(&self, b: &mut Builder) -> Lowering<Lowered> {
        let l = self.0.lower(b)?;
        Ok(Lowered { value: l.value.map(&mut |n| Ok(unary_node(U::OP, n, <V::T as DeviceElem>::IS_BOOL)))?, groups: l.groups })
    }
}
pub struct MapSym<V, F>(V, F);                                // returned by the new `View::map_sym(|x: Sym<T>| ...)`: traced ONCE
impl<V: Lower, U: DeviceElem, F: Fn(Sym<V::T>) -> Sym<U>> Lower for MapSym<V, F> where V::T: DeviceElem {
    fn lower(&self, b: &mut Builder) -> Lowering<Lowered> {
        let l = self.0.lower(b)?;
        Ok(Lowered { value: l.value.map(&mut |n| Ok((self.1)(Sym::of(n.clone())).node))?, groups: l.groups })
    }
}
/// `v.rows::<I, J>().map(fold)` with `fold: Fold<B, T>` (src/view.rs:617-622, 1341, 250-252): FOLD over J's axes, iterated
/// sequentially, last-fastest — the reference's only spelling of a reduction.
impl<V: Copy + Lower, I: Index, J: Index, B: DeviceBinary> Lower for Map<Rows<V, I, J>, Fold<B, V::T>> where V::T: DeviceElem {
    fn lower(&self, b: &mut Builder) -> Lowering<Lowered> {
        let l = (self.0).0.lower(b)?;
        let mut parts = split_groups(l.groups, &[I::LEAVES, J::LEAVES]);
        let (gj, gi) = (parts.pop().unwrap(), parts.pop().unwrap());
        let inner = match l.value { Value::Scalar(n) => n, Value::Tuple(_) => return Err(Unsupported("fold over tuple-typed elements".into())) };
        let mut f = Node::new(ffi::FOLD, inner.dtype);
        f.op = B::OP; f.children = vec![inner]; f.imm = self.1.init.bits(); f.red_axes = flat(&gj);
        Ok(Lowered { groups: gi, value: Value::Scalar(Rc::new(f)) })
    }
}

// ---- Compose (src/view.rs:897-912) and MapAxis (src/view.rs:1140-1170): every Array load under the source becomes a
//      bounds-checked GATHER whose index components are the index view's value trees ----------------------------------------
fn gather(node: &Rc<Node>, table: &HashMap<Axis, (Rc<Node>, u64)>, memo: &mut HashMap<*const Node, Rc<Node>>) -> Lowering<Rc<Node>> {
    if let Some(hit) = memo.get(&Rc::as_ptr(node)) { return Ok(hit.clone()); }
    let out = match node.kind {
        ffi::CONST => node.clone(),
        ffi::UNARY | ffi::BINARY => { let mut n = (**node).clone(); n.children = node.children.iter().map(|c| gather(c, table, memo)).collect::<Lowering<_>>()?; Rc::new(n) }
        ffi::CONCAT if !table.contains_key(&node.concat.unwrap().0) => {
            let mut n = (**node).clone(); n.children = node.children.iter().map(|c| gather(c, table, memo)).collect::<Lowering<_>>()?; Rc::new(n)
        }
        ffi::LEAF | ffi::GATHER => {
            let mut n = (**node).clone();
            n.children = node.children.iter().map(|c| gather(c, table, memo)).collect::<Lowering<_>>()?;
            n.stride.clear();
            for (a, s) in &node.stride {
                match table.get(a) {
                    Some((comp, bound)) => { n.children.push(comp.clone()); n.gstride.push(*s); n.bound.push(*bound); }
                    None => n.stride.push((*a, *s)),
                }
            }
            if n.children.is_empty() { node.clone() } else { n.kind = ffi::GATHER; Rc::new(n) }
        }
        ffi::IOTA => {
            let hit: Vec<Axis> = node.stride.iter().map(|(a, _)| *a).filter(|a| table.contains_key(a)).collect();
            if hit.is_empty() { node.clone() }
            else if node.stride.len() == 1 && node.stride[0].1 == 1 && node.offset == 0 { table[&hit[0]].0.clone() }   // All::at(index) = index (src/index.rs:185)
            else { return Err(Unsupported("compose onto a compound All".into())) }
        }
        _ => return Err(Unsupported("compose onto a view containing diagonal() or a fold".into())),
    };
    memo.insert(Rc::as_ptr(node), out.clone());
    Ok(out)
}
fn index_components(value: &Value, groups_w: &[Group]) -> Lowering<HashMap<Axis, (Rc<Node>, u64)>> {
    let comps = value.leaves();
    if comps.len() != groups_w.len() { return Err(Unsupported("compose: the index view's element type does not match the source's index type".into())); }
    let mut table = HashMap::new();
    for (node, g) in comps.iter().zip(groups_w) {
        if g.len() != 1 { return Err(Unsupported("gather through an index component whose axis has been re-split".into())); }
        table.insert(g[0], (cast_node(node, ffi::U64), g[0].length));
    }
    Ok(table)
}
impl<V: Lower, W: Lower<I = V::T>> Lower for Compose<V, W> {
    fn lower(&self, b: &mut Builder) -> Lowering<Lowered> {
        let (lv, lw) = (self.0.lower(b)?, self.1.lower(b)?);
        let table = index_components(&lv.value, &lw.groups)?;
        let mut memo = HashMap::new();
        Ok(Lowered { groups: lv.groups, value: lw.value.map(&mut |n| gather(n, &table, &mut memo))? })
    }
}
impl<V: Lower, I: Index, W: Lower, J: Index> Lower for MapAxis<V, I, W, J> where W::T: Index {
    fn lower(&self, b: &mut Builder) -> Lowering<Lowered> {
        let (lv, lw) = (self.0.lower(b)?, self.2.lower(b)?);
        let mut parts = split_groups(lv.groups, &[I::LEAVES, <W::T as Index>::LEAVES, J::LEAVES]);
        let (gj, gt, gi) = (parts.pop().unwrap(), parts.pop().unwrap(), parts.pop().unwrap());
        let table = index_components(&lw.value, &gt)?;
        let mut memo = HashMap::new();
        let mut groups = gi; groups.extend(lw.groups); groups.extend(gj);
        Ok(Lowered { groups, value: lv.value.map(&mut |n| gather(n, &table, &mut memo))? })
    }
}

// ---- Zip (src/view.rs:1178-1198) + Broadcast (src/broadcast.rs:22-162): axes unified leaf by leaf; an operand lacking an
//      axis simply has no stride on it.  `BroadcastLeaves` is the type-level walk of Broadcast::index over the leaf lists:
//      per result leaf, which operand(s) have it (added to the Broadcast trait as an associated const table). -------------
#[derive(Copy, Clone)] pub enum Side { Both, SelfOnly, OtherOnly }
// pub trait Broadcast<Other: Index>: Index { ...; /* NEW */ fn leaf_sides(out: &mut Vec<Side>); }
//   NonTuple vs itself: one `Both` per leaf; `()` vs J: J::LEAVES x OtherOnly; I vs `()`: I::LEAVES x SelfOnly; tuples: concatenation.
impl<V: Lower, W: Lower, B: DeviceBinary> Lower for Zip<V, W, B> where V::I: Broadcast<W::I>, B: Binary<V::T, W::T>, B::Output: Clone {
    fn lower(&self, b: &mut Builder) -> Lowering<Lowered> {
        let (lv, lw) = (self.0.lower(b)?, self.1.lower(b)?);
        let mut sides = vec![];
        <V::I as Broadcast<W::I>>::leaf_sides(&mut sides);
        let (mut gv, mut gw) = (lv.groups.into_iter(), lw.groups.into_iter());
        let (mut tv, mut tw, mut groups) = (HashMap::new(), HashMap::new(), vec![]);
        for s in sides {
            groups.push(match s {
                Side::SelfOnly => gv.next().unwrap(),
                Side::OtherOnly => gw.next().unwrap(),
                Side::Both => unify_group(b, &gv.next().unwrap(), &gw.next().unwrap(), &mut tv, &mut tw)?,   // sizes were checked by Zip::size ("Unequal sizes", src/broadcast.rs:38)
            });
        }
        let (vv, vw) = (subst_value(&lv.value, &tv)?, subst_value(&lw.value, &tw)?);
        match (vv, vw) {
            (Value::Scalar(x), Value::Scalar(y)) => Ok(Lowered { groups, value: Value::Scalar(binary_node(B::OP, &x, &y)) }),
            _ => Err(Unsupported("an arithmetic operator on tuple-typed elements".into())),
        }
    }
}
impl<V: Lower, W: Lower> Lower for Zip<V, W, Pair> where V::I: Broadcast<W::I> {   // ops::Pair (src/ops.rs:25-29): a tuple VALUE (structure of arrays)
    fn lower(&self, b: &mut Builder) -> Lowering<Lowered> {
        // identical to the arithmetic impl up to the substitution of both sides; the result keeps both values side by side
        let (lv, lw) = (self.0.lower(b)?, self.1.lower(b)?);
        let mut sides = vec![];
        <V::I as Broadcast<W::I>>::leaf_sides(&mut sides);
        let (mut gv, mut gw) = (lv.groups.into_iter(), lw.groups.into_iter());
        let (mut tv, mut tw, mut groups) = (HashMap::new(), HashMap::new(), vec![]);
        for s in sides {
            groups.push(match s {
                Side::SelfOnly => gv.next().unwrap(),
                Side::OtherOnly => gw.next().unwrap(),
                Side::Both => unify_group(b, &gv.next().unwrap(), &gw.next().unwrap(), &mut tv, &mut tw)?,
            });
        }
        Ok(Lowered { groups, value: Value::Tuple(vec![subst_value(&lv.value, &tv)?, subst_value(&lw.value, &tw)?]) })
    }
}

// ---- pure index remappings: no data movement, only axis bookkeeping ---------------------------------------------------------
impl<V: Lower, J: Index> Lower for Iso<V, J> {                 // src/view.rs:1238-1258: same leaf list, nothing to do
    fn lower(&self, b: &mut Builder) -> Lowering<Lowered> { self.0.lower(b) }
}
impl<V: Lower, I: Index> Lower for Coat<V, I> {                 // src/view.rs:1206-1230: the coated leaf owns all of V::I's axes
    fn lower(&self, b: &mut Builder) -> Lowering<Lowered> { let l = self.0.lower(b)?; Ok(Lowered { groups: vec![flat(&l.groups)], value: l.value }) }
}
impl<V: Lower, I: Index, X: Index, Y: Index, J: Index> Lower for Transpose<V, I, X, Y, J> {   // src/view.rs:1266-1294: (I,(Y,X),J) -> (I,(X,Y),J)
    fn lower(&self, b: &mut Builder) -> Lowering<Lowered> {
        let l = self.0.lower(b)?;
        let mut p = split_groups(l.groups, &[I::LEAVES, Y::LEAVES, X::LEAVES, J::LEAVES]);
        let (gj, gx, gy, gi) = (p.pop().unwrap(), p.pop().unwrap(), p.pop().unwrap(), p.pop().unwrap());
        let mut groups = gi; groups.extend(gx); groups.extend(gy); groups.extend(gj);
        Ok(Lowered { groups, value: l.value })
    }
}
impl<V: Lower, I: Index, J: Index> Lower for Row<V, I, J> {     // src/view.rs:1302-1322: I's axes become constants (offsets)
    fn lower(&self, b: &mut Builder) -> Lowering<Lowered> {
        let l = self.0.lower(b)?;
        let mut p = split_groups(l.groups, &[I::LEAVES, J::LEAVES]);
        let (gj, gi) = (p.pop().unwrap(), p.pop().unwrap());
        let isize = <(I::Size, J::Size)>::from_iso(self.0.size()).0;
        Ok(Lowered { groups: gj, value: subst_value(&l.value, &pin(&gi, self.1, isize))? })
    }
}
impl<V: Lower, I: Index, J: Index> Lower for Column<V, I, J> {  // src/view.rs:1350-1370
    fn lower(&self, b: &mut Builder) -> Lowering<Lowered> {
        let l = self.0.lower(b)?;
        let mut p = split_groups(l.groups, &[I::LEAVES, J::LEAVES]);
        let (gj, gi) = (p.pop().unwrap(), p.pop().unwrap());
        let jsize = <(I::Size, J::Size)>::from_iso(self.0.size()).1;
        Ok(Lowered { groups: gi, value: subst_value(&l.value, &pin(&gj, self.2, jsize))? })
    }
}
impl<V: Lower, I: Index, X: Index, J: Index> Lower for FromUsize<V, I, X, J> {   // src/view.rs:993-1021: the usize axis is re-split by X, row-major (:1019)
    fn lower(&self, b: &mut Builder) -> Lowering<Lowered> {
        let l = self.0.lower(b)?;
        let mut p = split_groups(l.groups, &[I::LEAVES, 1, J::LEAVES]);
        let (gj, gk, gi) = (p.pop().unwrap(), p.pop().unwrap(), p.pop().unwrap());
        let old = &gk[0];
        let mut lens_x = vec![];
        X::leaf_lengths(self.1, &mut lens_x);
        let flat_x: Vec<u64> = lens_x.iter().flatten().copied().collect();
        let (pieces, map_old, map_x) = lower::refine(&old.iter().map(|a| a.length).collect::<Vec<_>>(), &flat_x)?;
        let axes: Vec<Axis> = pieces.iter().map(|n| b.axis(*n)).collect();
        let table: HashMap<Axis, Sub> = old.iter().zip(&map_old).map(|(a, idx)| (*a, sub_of(&idx.iter().map(|k| axes[*k]).collect::<Vec<_>>()))).collect();
        let (mut gx, mut at) = (vec![], 0usize);
        for leaf in &lens_x { let mut g = vec![]; for _ in leaf { g.extend(map_x[at].iter().map(|k| axes[*k])); at += 1; } gx.push(g); }
        let mut groups = gi; groups.extend(gx); groups.extend(gj);
        Ok(Lowered { groups, value: subst_value(&l.value, &table)? })
    }
}
impl<V: Lower, I: Index, X: Index, J: Index> Lower for ToUsize<V, I, X, J> {     // src/view.rs:1029-1059: X's axes become ONE group (same run of positions)
    fn lower(&self, b: &mut Builder) -> Lowering<Lowered> {
        let l = self.0.lower(b)?;
        let mut p = split_groups(l.groups, &[I::LEAVES, X::LEAVES, J::LEAVES]);
        let (gj, gx, gi) = (p.pop().unwrap(), p.pop().unwrap(), p.pop().unwrap());
        let mut groups = gi; groups.push(flat(&gx)); groups.extend(gj);
        Ok(Lowered { groups, value: l.value })
    }
}
impl<V: Lower, I: Index, J: Index, K: Index> Lower for InsertOne<V, I, J, K> {   // src/view.rs:1067-1096: fresh length-1 axes
    fn lower(&self, b: &mut Builder) -> Lowering<Lowered> {
        let l = self.0.lower(b)?;
        let mut p = split_groups(l.groups, &[I::LEAVES, K::LEAVES]);
        let (gk, gi) = (p.pop().unwrap(), p.pop().unwrap());
        let mut lens = vec![];
        J::leaf_lengths(self.1, &mut lens);
        let gjn: Vec<Group> = lens.iter().map(|g| g.iter().map(|n| b.axis(*n)).collect()).collect();
        let mut groups = gi; groups.extend(gjn); groups.extend(gk);
        Ok(Lowered { groups, value: l.value })
    }
}
impl<V: Lower, I: Index, J: Index, K: Index> Lower for RemoveOne<V, I, J, K> {   // src/view.rs:1104-1132: J's axes pinned at 0
    fn lower(&self, b: &mut Builder) -> Lowering<Lowered> {
        let l = self.0.lower(b)?;
        let mut p = split_groups(l.groups, &[I::LEAVES, J::LEAVES, K::LEAVES]);
        let (gk, gj, gi) = (p.pop().unwrap(), p.pop().unwrap(), p.pop().unwrap());
        let table: HashMap<Axis, Sub> = flat(&gj).into_iter().map(|a| (a, Sub::pin(0))).collect();
        let mut groups = gi; groups.extend(gk);
        Ok(Lowered { groups, value: subst_value(&l.value, &table)? })
    }
}
impl<V: Lower, W: Lower, I: Index, J: Index> Lower for Concat<V, W, I, J> {      // src/view.rs:920-946: CONCAT over a fresh axis of length n_v + n_w
    fn lower(&self, b: &mut Builder) -> Lowering<Lowered> {
        let (lv, lw) = (self.0.lower(b)?, self.1.lower(b)?);
        let (mut pv, mut pw) = (split_groups(lv.groups, &[I::LEAVES, 1, J::LEAVES]), split_groups(lw.groups, &[I::LEAVES, 1, J::LEAVES]));
        let (vj, vk, vi) = (pv.pop().unwrap(), pv.pop().unwrap(), pv.pop().unwrap());
        let (wj, wk, wi) = (pw.pop().unwrap(), pw.pop().unwrap(), pw.pop().unwrap());
        if vk[0].len() != 1 || wk[0].len() != 1 { return Err(Unsupported("concat along an axis that is a merged group (to_usize) needs device div/mod".into())); }
        let (n_v, n_w) = (vk[0][0].length, wk[0][0].length);
        let k = b.axis(n_v + n_w);
        let (mut tv, mut tw) = (HashMap::new(), HashMap::new());
        let gi: Vec<Group> = vi.iter().zip(&wi).map(|(x, y)| unify_group(b, x, y, &mut tv, &mut tw)).collect::<Lowering<_>>()?;
        let gj: Vec<Group> = vj.iter().zip(&wj).map(|(x, y)| unify_group(b, x, y, &mut tv, &mut tw)).collect::<Lowering<_>>()?;
        tv.insert(vk[0][0], Sub::rename(k));
        tw.insert(wk[0][0], Sub { constant: -(n_v as i64), terms: vec![(k, 1)] });   // W is addressed with k - len(V) (:943)
        let (vv, vw) = (subst_value(&lv.value, &tv)?, subst_value(&lw.value, &tw)?);
        let nodes: Vec<Value> = vv.leaves().iter().zip(vw.leaves().iter()).map(|(x, y)| {
            let mut c = Node::new(ffi::CONCAT, x.dtype);
            c.children = vec![x.clone(), y.clone()]; c.concat = Some((k, n_v));
            Value::Scalar(Rc::new(c))
        }).collect();
        let mut groups = gi; groups.push(vec![k]); groups.extend(gj);
        Ok(Lowered { groups, value: if nodes.len() == 1 { nodes.into_iter().next().unwrap() } else { Value::Tuple(nodes) } })
    }
}
// Not lowerable (the default `from_view` path keeps serving them on the host): `Map` with an opaque closure, `Nested` /
// `nested_collect` (views as elements), `Rows` / `Columns` on their own, `ViewRef` / `ViewMut` (in-place writes).
