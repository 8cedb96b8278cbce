// ---- append to src/index.rs (and src/int.rs for usize / Fixed / Reversed) --------------------------------------------
// UNVERIFIED — see rust/README.md.  Mirrors multidimension_b200/index.py::{type_leaves, leaf_lengths, index_positions}.
//
// The ONLY run-time content of an index type on the device is the list of its position-space axes: flatten the type to
// its NonTuple leaves (src/tuple.rs:60-176); `Index::to_usize` is row-major over them (src/index.rs:109-114) and
// `Index::each` walks them last-fastest (src/index.rs:122-124).  Three provided-by-impl items are added to the trait:
//
//   pub trait Index: Debug + Copy + PartialEq {
//       ...
//       /// Number of NonTuple leaves of the flattened type (`()` has none, `Coated<I>` is ONE leaf).
//       const LEAVES: usize;
//       /// Per leaf, the lengths of its position axes (one axis, except `Coated<I>`: all of I's).
//       fn leaf_lengths(size: Self::Size, out: &mut Vec<Vec<u64>>);
//       /// Per leaf, the positions of `self` along those axes, with the reference's bounds assert (src/int.rs:16-19).
//       fn leaf_positions(self, size: Self::Size, out: &mut Vec<Vec<u64>>);
//   }

impl Index for usize {                                     // src/int.rs:9-26
    // ... existing items ...
    const LEAVES: usize = 1;
    fn leaf_lengths(size: usize, out: &mut Vec<Vec<u64>>) { out.push(vec![size as u64]); }
    fn leaf_positions(self, size: usize, out: &mut Vec<Vec<u64>>) {
        assert!(self < size, "Index {:?} is out of bounds for size {:?}", self, size);   // src/int.rs:17
        out.push(vec![self as u64]);
    }
}

impl<const N: usize> Index for Fixed<N> {                   // src/int.rs:33-54 (StaticIndex)
    const LEAVES: usize = 1;
    fn leaf_lengths(_: (), out: &mut Vec<Vec<u64>>) { out.push(vec![N as u64]); }
    fn leaf_positions(self, _: (), out: &mut Vec<Vec<u64>>) {
        assert!(self.0 < N, "Index {:?} is out of bounds for size {:?}", self.0, N);   // the slice access of src/array.rs:86
        out.push(vec![self.0 as u64]);
    }
}

impl Index for Reversed {                                   // src/int.rs:58-86: position p <-> index size-1-p
    const LEAVES: usize = 1;
    fn leaf_lengths(size: usize, out: &mut Vec<Vec<u64>>) { out.push(vec![size as u64]); }
    fn leaf_positions(self, size: usize, out: &mut Vec<Vec<u64>>) { out.push(vec![((size - 1) - self.0) as u64]); }
}

impl Index for () {                                         // src/index.rs:230-234
    const LEAVES: usize = 0;
    fn leaf_lengths(_: (), _out: &mut Vec<Vec<u64>>) {}
    fn leaf_positions(self, _: (), _out: &mut Vec<Vec<u64>>) {}
}

impl Index for bool {                                       // src/index.rs:236-240
    const LEAVES: usize = 1;
    fn leaf_lengths(_: (), out: &mut Vec<Vec<u64>>) { out.push(vec![2]); }
    fn leaf_positions(self, _: (), out: &mut Vec<Vec<u64>>) { out.push(vec![self as u64]); }
}

impl<I: Index> Index for Option<I> {                        // src/index.rs:244-276: None is position 0, Some(i) is 1 + i.to_usize()
    const LEAVES: usize = 1;
    fn leaf_lengths(size: I::Size, out: &mut Vec<Vec<u64>>) { out.push(vec![1 + I::length(size) as u64]); }
    fn leaf_positions(self, size: I::Size, out: &mut Vec<Vec<u64>>) {
        out.push(vec![match self { None => 0, Some(i) => 1 + i.to_usize(size) as u64 }]);
    }
}

impl<I: Index> Index for Coated<I> {                        // src/index.rs:156-171: hides I's tuple structure -> one leaf
    const LEAVES: usize = 1;
    fn leaf_lengths(size: Coated<I::Size>, out: &mut Vec<Vec<u64>>) {
        let mut inner = vec![];
        I::leaf_lengths(size.0, &mut inner);
        out.push(inner.into_iter().flatten().collect());
    }
    fn leaf_positions(self, size: Coated<I::Size>, out: &mut Vec<Vec<u64>>) {
        let mut inner = vec![];
        self.0.leaf_positions(size.0, &mut inner);
        out.push(inner.into_iter().flatten().collect());
    }
}

macro_rules! leaves_for_tuple {                             // src/index.rs:75-154
    ($(($($T:ident $i:tt),+))+) => { $(
        impl<$($T: Index),+> Index for ($($T,)+) {
            const LEAVES: usize = 0 $(+ $T::LEAVES)+;
            fn leaf_lengths(size: Self::Size, out: &mut Vec<Vec<u64>>) { $($T::leaf_lengths(size.$i, out);)+ }
            fn leaf_positions(self, size: Self::Size, out: &mut Vec<Vec<u64>>) { $(self.$i.leaf_positions(size.$i, out);)+ }
        }
    )+ };
}
leaves_for_tuple! { (A 0) (A 0, B 1) (A 0, B 1, C 2) }
