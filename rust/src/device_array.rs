//! `DeviceArray<I, T>`: the reference's `Array<I, T>` (src/array.rs:5-114) with its boxed buffer in HBM.
//! UNVERIFIED — see rust/README.md.  Mirrors multidimension_b200/runtime.py + view.py::Array / View.collect.
use std::cell::RefCell;
use std::marker::PhantomData;
use std::rc::Rc;

use crate::device_ops::DeviceElem;
use crate::ffi;
use crate::lower::{self, Axis, Builder, Lower, Lowered, Lowering, Node, Peers, Value};
use crate::{Index, Isomorphic, NewView, Push, View};

/// mdim_ctx: one per GPU per process, single-threaded like the reference (no Send / Sync story of its own).
pub struct Context { raw: *mut ffi::MdimCtx }

impl Context {
    /// Panics when the library finds no sm_100 device: there is no CPU fallback behind the C ABI.
    pub fn new(device: i32) -> Context {
        let mut raw = core::ptr::null_mut();
        let st = unsafe { ffi::mdim_init(device, &mut raw) };
        if st != ffi::MDIM_OK { panic!("mdim_init(device = {}): status {}", device, st) }
        Context { raw }
    }
    /// Turns a non-OK status into the reference's panic (message text from mdim_error_info, e.g.
    /// "Index 7 is out of bounds for size 3", src/int.rs:17).  Errors never unwind across the C boundary.
    pub fn check(&self, st: i32) {
        if st == ffi::MDIM_OK { return }
        let mut info: ffi::MdimErrorInfo = unsafe { core::mem::zeroed() };
        unsafe { ffi::mdim_last_error(self.raw, &mut info) };
        let msg = unsafe { std::ffi::CStr::from_ptr(info.message.as_ptr()) }.to_string_lossy().into_owned();
        panic!("{}", if msg.is_empty() { format!("mdim status {}", st) } else { msg });
    }
    pub fn raw(&self) -> *mut ffi::MdimCtx { self.raw }
}
impl Drop for Context { fn drop(&mut self) { unsafe { ffi::mdim_shutdown(self.raw); } } }

thread_local! { static CTX: RefCell<Option<Rc<Context>>> = RefCell::new(None); }
/// The context of this thread (LOCAL_RANK picks the GPU, as torchrun / mpirun export it).
pub fn context() -> Rc<Context> {
    CTX.with(|c| c.borrow_mut().get_or_insert_with(|| {
        let dev = std::env::var("LOCAL_RANK").ok().and_then(|s| s.parse().ok()).unwrap_or(0);
        Rc::new(Context::new(dev))
    }).clone())
}

/// The device-resident `Box<[T]>` (src/array.rs:5-8): freed on Drop, never aliased with an output.
struct DeviceBox { ptr: *mut core::ffi::c_void, len: usize, ctx: Rc<Context> }
impl Drop for DeviceBox { fn drop(&mut self) { unsafe { ffi::mdim_buf_free(self.ctx.raw(), self.ptr); } } }

pub struct DeviceArray<I: Index, T: DeviceElem> {
    size: I::Size,
    items: Rc<DeviceBox>,
    /// Set when the Array is sharded over the GPUs of the box (equal blocks, block p on rank p): mdim_node.peer[].
    peers: Option<Peers>,
    _t: PhantomData<T>,
}

impl<I: Index, T: DeviceElem> DeviceArray<I, T> {
    fn uninit(size: I::Size) -> Self {
        let ctx = context();
        let len = I::length(size);
        let mut ptr = core::ptr::null_mut();
        ctx.check(unsafe { ffi::mdim_buf_alloc(ctx.raw(), len * core::mem::size_of::<T>(), &mut ptr) });
        DeviceArray { size, items: Rc::new(DeviceBox { ptr, len, ctx }), peers: None, _t: PhantomData }
    }
    /// `Array::new` (src/array.rs:28-30) + upload; panics like `new_inner` (src/array.rs:11-14) on a length mismatch.
    pub fn new(size: impl Isomorphic<I::Size>, items: impl AsRef<[T]>) -> Self {
        let size = size.to_iso();
        let items = items.as_ref();
        assert_eq!(I::length(size), items.len());
        let a = Self::uninit(size);
        a.items.ctx.check(unsafe { ffi::mdim_upload(a.items.ctx.raw(), a.items.ptr, items.as_ptr() as *const _, items.len() * core::mem::size_of::<T>()) });
        a
    }
    /// `Array::to_raw` (src/array.rs:54): downloads.
    pub fn to_raw(&self) -> Box<[T]> {
        let mut v: Vec<T> = Vec::with_capacity(self.items.len);
        self.items.ctx.check(unsafe { ffi::mdim_download(self.items.ctx.raw(), v.as_mut_ptr() as *mut _, self.items.ptr, self.items.len * core::mem::size_of::<T>()) });
        unsafe { v.set_len(self.items.len) };
        v.into_boxed_slice()
    }
    /// `Array::iso` (src/array.rs:57-62): no data movement.
    pub fn iso<J: Index>(self) -> DeviceArray<J, T> where J::Size: Isomorphic<I::Size> {
        DeviceArray { size: J::Size::from_iso(self.size), items: self.items, peers: self.peers, _t: PhantomData }
    }
}

impl<I: Index, T: DeviceElem> View for DeviceArray<I, T> {
    type I = I;
    type T = T;
    fn size(&self) -> I::Size { self.size }
    /// src/array.rs:81: a host-side probe (downloads one element, bounds-checked like `Index::to_usize`).
    fn at(&self, index: I) -> T {
        let k = index.to_usize(self.size);
        assert!(k < self.items.len);
        let mut one = core::mem::MaybeUninit::<T>::uninit();
        self.items.ctx.check(unsafe { ffi::mdim_download(self.items.ctx.raw(), one.as_mut_ptr() as *mut _,
                                                         (self.items.ptr as *const u8).add(k * core::mem::size_of::<T>()) as *const _, core::mem::size_of::<T>()) });
        unsafe { one.assume_init() }
    }
}

impl<I: Index, T: DeviceElem> Lower for DeviceArray<I, T> {
    /// LEAF: base pointer + row-major strides over fresh named axes (`Index::to_usize`, src/index.rs:109-114).
    fn lower(&self, b: &mut Builder) -> Lowering<Lowered> {
        let mut lens = vec![];
        I::leaf_lengths(self.size, &mut lens);
        let groups: Vec<Vec<Axis>> = lens.iter().map(|g| g.iter().map(|n| b.axis(*n)).collect()).collect();
        let mut n = Node::new(ffi::LEAF, T::DTYPE);
        n.data = self.items.ptr;
        n.peers = self.peers.clone();
        let mut acc = 1i64;
        for a in groups.iter().flatten().rev() { n.stride.push((*a, acc)); acc *= a.length as i64; }
        Ok(Lowered { groups, value: Value::Scalar(Rc::new(n)) })
    }
}

/// The sink side: `collect::<DeviceArray<_, _>>()`.  `NewView::from_view` is the ONE hook `View::collect`
/// (src/view.rs:146-150) needs — see rust/patches/view_rs.rs for the trait change; the element-at-a-time
/// `Push` protocol (src/view.rs:8-38) stays the default for every other sink.
impl<I: Index, T: DeviceElem> NewView for DeviceArray<I, T> {
    type Buffer = Vec<T>;
    fn new_view(size: I::Size, callback: impl FnOnce(&mut Vec<T>)) -> Self {
        let mut buffer = Vec::with_capacity(I::length(size));   // a host-computed view still lands on the device
        callback(&mut buffer);
        DeviceArray::new(size, buffer)
    }
    fn from_view<V: View<I = I, T = T> + Lower>(v: &V) -> Self {
        let mut b = Builder::new();
        let lowered = v.lower(&mut b).unwrap_or_else(|e| panic!("not lowerable to the device: {}", e.0));
        let axes: Vec<Axis> = lowered.groups.iter().flatten().copied().collect();
        let root = match &lowered.value { Value::Scalar(n) => n.clone(), Value::Tuple(_) => panic!("tuple-typed elements: collect the components (structure of arrays)") };
        let em = lower::emit(&root, &axes).unwrap_or_else(|e| panic!("not lowerable to the device: {}", e.0));
        let out = DeviceArray::uninit(v.size());
        out.items.ctx.check(unsafe { ffi::mdim_collect(out.items.ctx.raw(), &em.expr, out.items.ptr, 0) });
        out
    }
}
