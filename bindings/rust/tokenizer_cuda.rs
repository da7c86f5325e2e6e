// Sketch of the feature-gated bodies a gtars maintainer would add (WRITTEN, NOT BUILT — no Rust toolchain here).
// Public signatures stay byte-for-byte what they are today; `--features cuda` switches the BODY to the FFI path.
//
//   gtars-tokenizers/src/tokenizer.rs:165-171   Tokenizer::encode
//   gtars-overlaprs/src/multi_chrom_overlapper.rs:483-498   MultiChromOverlapper::count_overlaps
//   gtars-igd/src/igd.rs:563-590   Igd::count_region_hits
#[cfg(feature = "cuda")]
mod cuda {
    use gtars_overlaprs_sys as sys;
    use std::{collections::HashMap, ffi::CStr, ptr, slice};

    fn check(status: i32) -> Result<(), String> {
        if status == sys::GTGPU_OK { return Ok(()); }
        Err(unsafe { CStr::from_ptr(sys::gtgpu_last_error()) }.to_string_lossy().into_owned())
    }

    /// Built once next to `Tokenizer::core` (create_tokenize_core_from_universe, utils/mod.rs:49-99): the same
    /// per-chromosome interval lists, flattened; val = region_to_id[id_to_region[val]] so ids need no string round trip.
    pub struct GpuCore { pub index: *mut sys::gtgpu_index, pub chrom_ids: HashMap<String, u32>, pub unk_id: u32 }

    impl super::Tokenizer {
        pub fn encode(&self, regions: &[Region]) -> Result<Vec<u32>, TokenizerError> {
            let g = &self.gpu;                                    // GpuCore, built in Tokenizer::new under the feature
            let (mut chr, mut start, mut end) = (Vec::with_capacity(regions.len()), Vec::new(), Vec::new());
            for r in regions {
                chr.push(*g.chrom_ids.get(&r.chr).unwrap_or(&sys::GTGPU_UNKNOWN_CHROM));   // tokenizer.rs:144
                start.push(r.start);
                end.push(r.end);
            }
            let file_offsets = [0u64, regions.len() as u64];     // one encode() call = one "file"
            let mut tok_offsets = [0u64; 2];
            let mut buf = ptr::null_mut();
            check(unsafe { sys::gtgpu_tokenize_files(g.index, 1, file_offsets.as_ptr(), chr.as_ptr(), start.as_ptr(),
                          end.as_ptr(), g.unk_id, tok_offsets.as_mut_ptr(), &mut buf) }).map_err(TokenizerError::Gpu)?;
            let ids = unsafe { slice::from_raw_parts(sys::gtgpu_buf_data(buf) as *const u32, sys::gtgpu_buf_len(buf) as usize) }.to_vec();
            unsafe { sys::gtgpu_buf_free(buf) };
            Ok(ids)                                               // [unk] rule already applied (tokenizer.rs:158-160)
        }
    }

    impl<I, T> super::MultiChromOverlapper<I, T> {
        pub fn count_overlaps(&self, query: &RegionSet, min_overlap: Option<i32>) -> Vec<usize> {
            let (chr, start, end) = self.flatten(query);          // same chrom-id map as above, unknown -> u32::MAX
            let mut counts = vec![0u32; chr.len()];
            check(unsafe { sys::gtgpu_count(self.gpu_index, chr.len() as u64, chr.as_ptr(), start.as_ptr(), end.as_ptr(),
                          min_overlap.unwrap_or(0), counts.as_mut_ptr()) }).expect("gtgpu_count");
            counts.into_iter().map(|c| c as usize).collect()
        }
    }

    impl super::Igd {
        pub fn count_region_hits(&self, regions: &RegionSet, min_overlap: i32) -> Vec<u64> {
            let (chr, start, end) = self.flatten(regions);
            let set_offsets = [0u64, chr.len() as u64];
            let mut totals = vec![0u64; self.file_info.len()];
            check(unsafe { sys::gtgpu_igd_count_region_hits(self.gpu_igd, 1, set_offsets.as_ptr(), chr.as_ptr(), start.as_ptr(),
                          end.as_ptr(), min_overlap, totals.as_mut_ptr()) }).expect("gtgpu_igd_count_region_hits");
            totals
        }
    }
}
