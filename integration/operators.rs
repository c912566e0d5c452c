// Sketch (not compiled here: no Rust toolchain): the FFI-backed operator twins of INTEGRATION.md section 3 over daft_b200_sys.rs.
// Paths in comments are the reference's: each impl replaces the named operator when the runner is "b200".
use std::ffi::{CStr, CString};
use std::os::raw::c_char;
use std::ptr;

use common_error::{DaftError, DaftResult};
use daft_b200_sys as sys;
use daft_recordbatch::RecordBatch;

/// status + malloc'd message -> DaftError (the daft-ext loader does the same, src/daft-ext-internal/src/module.rs)
unsafe fn check(rc: i32, err: *mut c_char) -> DaftResult<()> {
    if rc == 0 {
        return Ok(());
    }
    let msg = CStr::from_ptr(err).to_string_lossy().into_owned();
    sys::b200_free_string(err);
    Err(DaftError::ComputeError(msg))
}

/// A device-resident batch; operators pass these between each other, so a pipeline stays in HBM.
pub struct Batch(*mut sys::B200Batch);
unsafe impl Send for Batch {}
unsafe impl Sync for Batch {} // read-only use from several threads is part of the library's contract
impl Drop for Batch {
    fn drop(&mut self) {
        unsafe { sys::b200_batch_free(self.0) }
    }
}

impl Batch {
    /// H2D: RecordBatch::to_arrow -> one struct ArrowArray (what daft-ext already produces, src/daft-ext-internal/src/function.rs)
    pub fn from_record_batch(rb: &RecordBatch) -> DaftResult<Self> {
        let (mut array, mut schema) = arrow::ffi::to_ffi(&rb.to_arrow_struct_array()?.into_data())?;
        let (mut out, mut err) = (ptr::null_mut(), ptr::null_mut());
        unsafe { check(sys::b200_batch_import(&mut array, &mut schema, &mut out, &mut err), err)? };
        Ok(Self(out))
    }
    /// D2H (waits for the library stream)
    pub fn to_record_batch(&self) -> DaftResult<RecordBatch> {
        let (mut array, mut schema) = (arrow::ffi::FFI_ArrowArray::empty(), arrow::ffi::FFI_ArrowSchema::empty());
        let mut err = ptr::null_mut();
        unsafe { check(sys::b200_batch_export(self.0, &mut array, &mut schema, &mut err), err)? };
        RecordBatch::from_arrow_struct_array(unsafe { arrow::ffi::from_ffi(array, &schema)? })
    }
    fn call_plan(&self, f: unsafe extern "C" fn(*const sys::B200Batch, *const u8, usize, *mut *mut sys::B200Batch, *mut *mut c_char) -> i32,
                 plan: &[u8]) -> DaftResult<Self> {
        let (mut out, mut err) = (ptr::null_mut(), ptr::null_mut());
        unsafe { check(f(self.0, plan.as_ptr(), plan.len(), &mut out, &mut err), err)? };
        Ok(Self(out))
    }
}

/// Project / Filter: IntermediateOperator::execute (src/daft-local-execution/src/intermediate_ops/intermediate_op.rs:42-71).
/// `plan` = the node's bound expressions serialised once at construction.
pub struct B200Project { plan: Vec<u8> }
impl B200Project {
    pub fn execute(&self, input: &Batch) -> DaftResult<Batch> { input.call_plan(sys::b200_eval_expression_list, &self.plan) }
}
pub struct B200Filter { plan: Vec<u8> }
impl B200Filter {
    pub fn execute(&self, input: &Batch) -> DaftResult<Batch> { input.call_plan(sys::b200_filter, &self.plan) }
}

/// HashAggregate: BlockingSink::{make_state, sink, finalize} (sinks/grouped_aggregate.rs:146-488).  The three plans are the lists
/// populate_aggregation_stages_bound already computes (src/daft-local-plan/src/agg.rs:13-25).
pub struct B200GroupedAggregate(*mut sys::B200GroupbyState);
impl B200GroupedAggregate {
    pub fn make_state(partial: &[u8], fin: &[u8], project: &[u8]) -> DaftResult<Self> {
        let (mut out, mut err) = (ptr::null_mut(), ptr::null_mut());
        unsafe {
            check(sys::b200_groupby_create(partial.as_ptr(), partial.len(), fin.as_ptr(), fin.len(), project.as_ptr(), project.len(), &mut out, &mut err), err)?
        };
        Ok(Self(out))
    }
    pub fn sink(&self, input: &Batch) -> DaftResult<()> {
        let mut err = ptr::null_mut();
        unsafe { check(sys::b200_groupby_push(self.0, input.0, &mut err), err) }
    }
    pub fn finalize(&self) -> DaftResult<Batch> {
        let (mut out, mut err) = (ptr::null_mut(), ptr::null_mut());
        unsafe { check(sys::b200_groupby_finalize(self.0, &mut out, &mut err), err)? };
        Ok(Batch(out))
    }
}
impl Drop for B200GroupedAggregate {
    fn drop(&mut self) {
        unsafe { sys::b200_groupby_free(self.0) }
    }
}

/// HashJoin: JoinOperator::{build, finalize_build, probe, finalize_probe} (join/join_operator.rs:26-97, join/hash_join.rs:33-362).
pub struct B200ProbeTable(*mut sys::B200ProbeTable);
unsafe impl Send for B200ProbeTable {}
unsafe impl Sync for B200ProbeTable {} // Arc<dyn Probeable> in the reference: built once, probed from many tasks
impl B200ProbeTable {
    pub fn build(batches: &[&Batch], key_cols: &[i32], null_equals_null: &[u8], track_indices: bool) -> DaftResult<Self> {
        let (mut pt, mut err) = (ptr::null_mut(), ptr::null_mut());
        unsafe {
            check(sys::b200_probe_table_create(key_cols.len() as i32, null_equals_null.as_ptr(), track_indices as i32, &mut pt, &mut err), err)?;
            for b in batches {
                check(sys::b200_probe_table_add(pt, b.0, key_cols.as_ptr(), &mut err), err)?;
            }
            check(sys::b200_probe_table_build(pt, &mut err), err)?;
        }
        Ok(Self(pt))
    }
    /// inner / semi / anti probes need no state; left / right / outer pass the per-join matched-row bitmap
    pub fn probe(&self, input: &Batch, key_cols: &[i32], join_type: i32, build_on_left: bool) -> DaftResult<Batch> {
        let (mut out, mut err) = (ptr::null_mut(), ptr::null_mut());
        unsafe {
            check(sys::b200_hash_join_probe(self.0, ptr::null_mut(), input.0, key_cols.as_ptr(), join_type, build_on_left as i32, &mut out, &mut err), err)?
        };
        Ok(Batch(out)) // views: row ids over the inputs, gathered when an operator first reads a column
    }
}
impl Drop for B200ProbeTable {
    fn drop(&mut self) {
        unsafe { sys::b200_probe_table_free(self.0) }
    }
}

/// Scan task source for a local Parquet file (sources/scan_task.rs:33): one task per row-group range, pushed-down columns by name.
pub fn parquet_scan_task(path: &str, columns: &[&str], row_groups: std::ops::Range<i64>) -> DaftResult<Batch> {
    let cpath = CString::new(path).unwrap();
    let cols: Vec<CString> = columns.iter().map(|c| CString::new(*c).unwrap()).collect();
    let ptrs: Vec<*const c_char> = cols.iter().map(|c| c.as_ptr()).collect();
    let (mut out, mut err) = (ptr::null_mut(), ptr::null_mut());
    unsafe {
        check(sys::b200_parquet_read(cpath.as_ptr(), ptrs.as_ptr(), ptrs.len() as i32, row_groups.start, row_groups.end, &mut out, &mut err), err)?
    };
    Ok(Batch(out))
}
