//! `CudaProfiles`: the batched, B200-resident counterpart of zoe's `SharedProfiles`
//! (src/alignment/profile_set.rs:552-560).  Element-wise equal to
//! `SharedProfiles::sw_score_from_i8` / `sw_align_from_i8` (profile_set.rs:71-78, 136-145).
use std::ffi::CStr;

use zoe::{
    alignment::{Alignment, AlignmentStates, MaybeAligned, ProfileError, ScoreAndRanges, SeqSrc},
    data::{cigar::Ciglet, matrices::WeightMatrix},
};
use zoe_cuda_sys as sys;

pub struct CudaProfiles {
    ctx:               *mut sys::zoe_cuda_ctx,
    profiled_lens:     Vec<usize>,
    profiled_is_query: bool,
}

// One context per host thread (like LocalProfiles); it may be moved, not shared.
unsafe impl Send for CudaProfiles {}

#[derive(Debug)]
pub enum CudaError {
    Profile(ProfileError),
    Cuda { code: i32, message: String },
}

impl CudaProfiles {
    /// `targets` are the profiled sequences.  `profiled_is_query = false` means they are references and the
    /// sequences passed later are queries (`SeqSrc::Query`), the usual read-vs-reference shape.
    pub fn new_with_w256<const S: usize>(
        targets: &[&[u8]], matrix: &WeightMatrix<i8, S>, gap_open: i8, gap_extend: i8, profiled_is_query: bool,
        devices: &[i32],
    ) -> Result<Self, CudaError> {
        Self::new(targets, matrix, gap_open, gap_extend, (32, 16, 8), profiled_is_query, devices)
    }

    pub fn new_with_w128<const S: usize>(
        targets: &[&[u8]], matrix: &WeightMatrix<i8, S>, gap_open: i8, gap_extend: i8, profiled_is_query: bool,
        devices: &[i32],
    ) -> Result<Self, CudaError> {
        Self::new(targets, matrix, gap_open, gap_extend, (16, 8, 4), profiled_is_query, devices)
    }

    pub fn new_with_w512<const S: usize>(
        targets: &[&[u8]], matrix: &WeightMatrix<i8, S>, gap_open: i8, gap_extend: i8, profiled_is_query: bool,
        devices: &[i32],
    ) -> Result<Self, CudaError> {
        Self::new(targets, matrix, gap_open, gap_extend, (64, 32, 16), profiled_is_query, devices)
    }

    fn new<const S: usize>(
        targets: &[&[u8]], matrix: &WeightMatrix<i8, S>, gap_open: i8, gap_extend: i8, lanes: (i32, i32, i32),
        profiled_is_query: bool, devices: &[i32],
    ) -> Result<Self, CudaError> {
        let mut ctx = std::ptr::null_mut();
        let rc = unsafe { sys::zoe_cuda_create(&mut ctx, devices.as_ptr(), devices.len() as i32) };
        if rc != 0 {
            return Err(CudaError::Cuda { code: rc, message: "no usable CUDA device".into() });
        }
        let this = CudaProfiles { ctx, profiled_lens: targets.iter().map(|t| t.len()).collect(), profiled_is_query };
        let weights: Vec<i8> = matrix.weights.iter().flatten().copied().collect();
        let mut lut = [0u8; 256];
        for (b, slot) in lut.iter_mut().enumerate() {
            *slot = matrix.mapping.to_index(b as u8) as u8;
        }
        this.check(unsafe {
            sys::zoe_cuda_set_scoring(ctx, weights.as_ptr(), S as i32, lut.as_ptr(), gap_open, gap_extend, profiled_is_query as i32)
        }, gap_open, gap_extend)?;
        this.check(unsafe { sys::zoe_cuda_set_lanes(ctx, lanes.0, lanes.1, lanes.2) }, gap_open, gap_extend)?;
        let (concat, offsets) = pack(targets);
        this.check(unsafe { sys::zoe_cuda_set_profiled(ctx, concat.as_ptr(), offsets.as_ptr(), targets.len() as u32) }, gap_open, gap_extend)?;
        Ok(this)
    }

    /// `out[i * n_profiled + j] == profiles[j].sw_score_from_i8(seqs[i])`
    pub fn sw_score_batch(&self, seqs: &[&[u8]]) -> Result<Vec<MaybeAligned<u32>>, CudaError> {
        let (concat, offsets) = pack(seqs);
        let pairs = seqs.len() * self.profiled_lens.len();
        let (mut score, mut status, mut tier) = (vec![0u32; pairs], vec![0u8; pairs], vec![0u8; pairs]);
        self.check(unsafe {
            sys::zoe_cuda_sw_score_batch(self.ctx, concat.as_ptr(), offsets.as_ptr(), seqs.len() as u64,
                                         score.as_mut_ptr(), status.as_mut_ptr(), tier.as_mut_ptr())
        }, 0, 0)?;
        Ok(score.iter().zip(&status).map(|(&s, &st)| maybe(st, s)).collect())
    }

    /// `out[i * n_profiled + j] == profiles[j].sw_align_from_i8(seq_src(seqs[i]))` where `seq_src` is
    /// `SeqSrc::Query` when the profiled sequences are references, `SeqSrc::Reference` otherwise.
    pub fn sw_align_batch(&self, seqs: &[&[u8]]) -> Result<Vec<MaybeAligned<Alignment<u32>>>, CudaError> {
        self.align(seqs, false)
    }

    /// `out[i * n_profiled + j] == profiles[j].sw_align_from_i8_3pass(seq_src(seqs[i]))` (profile_set.rs:213-231):
    /// ranges from two score passes, then a banded alignment of the bounding box (three_pass.rs:21-104).
    pub fn sw_align_3pass_batch(&self, seqs: &[&[u8]]) -> Result<Vec<MaybeAligned<Alignment<u32>>>, CudaError> {
        self.align(seqs, true)
    }

    fn align(&self, seqs: &[&[u8]], three_pass: bool) -> Result<Vec<MaybeAligned<Alignment<u32>>>, CudaError> {
        let (concat, offsets) = pack(seqs);
        let np = self.profiled_lens.len();
        let pairs = seqs.len() * np;
        let mut score = vec![0u32; pairs];
        let (mut status, mut tier) = (vec![0u8; pairs], vec![0u8; pairs]);
        let (mut rs, mut re, mut qs, mut qe) = (vec![0u32; pairs], vec![0u32; pairs], vec![0u32; pairs], vec![0u32; pairs]);
        let mut coff = vec![0u64; pairs + 1];
        let mut cigar = vec![0u32; 16 * pairs + 1024];
        loop {
            let rc = unsafe {
                if three_pass {
                    sys::zoe_cuda_sw_align_3pass_batch(self.ctx, concat.as_ptr(), offsets.as_ptr(), seqs.len() as u64,
                        score.as_mut_ptr(), status.as_mut_ptr(), tier.as_mut_ptr(), rs.as_mut_ptr(), re.as_mut_ptr(),
                        qs.as_mut_ptr(), qe.as_mut_ptr(), cigar.as_mut_ptr(), coff.as_mut_ptr(), cigar.len() as u64)
                } else {
                    sys::zoe_cuda_sw_align_batch(self.ctx, concat.as_ptr(), offsets.as_ptr(), seqs.len() as u64,
                        score.as_mut_ptr(), status.as_mut_ptr(), tier.as_mut_ptr(), rs.as_mut_ptr(), re.as_mut_ptr(),
                        qs.as_mut_ptr(), qe.as_mut_ptr(), cigar.as_mut_ptr(), coff.as_mut_ptr(), cigar.len() as u64,
                        std::ptr::null_mut())
                }
            };
            if rc == sys::ZOE_CUDA_E_CIGAR_CAP {
                cigar.resize(coff[0] as usize + 16, 0);
                continue;
            }
            self.check(rc, 0, 0)?;
            break;
        }
        const OPS: [u8; 5] = [b'M', b'I', b'D', b'?', b'S'];
        let mut out = Vec::with_capacity(pairs);
        for (i, seq) in seqs.iter().enumerate() {
            for j in 0..np {
                let k = i * np + j;
                if status[k] != sys::ZOE_CUDA_SOME {
                    out.push(if status[k] == sys::ZOE_CUDA_OVERFLOWED { MaybeAligned::Overflowed } else { MaybeAligned::Unmapped });
                    continue;
                }
                let ciglets = cigar[coff[k] as usize..coff[k + 1] as usize]
                    .iter()
                    .map(|w| Ciglet { inc: (w >> 4) as usize, op: OPS[(w & 7) as usize] });
                let (ref_len, query_len) =
                    if self.profiled_is_query { (seq.len(), self.profiled_lens[j]) } else { (self.profiled_lens[j], seq.len()) };
                out.push(MaybeAligned::Some(Alignment {
                    score: score[k],
                    ref_range: rs[k] as usize..re[k] as usize,
                    query_range: qs[k] as usize..qe[k] as usize,
                    states: AlignmentStates::from_ciglets_unchecked(ciglets),
                    ref_len,
                    query_len,
                }));
            }
        }
        Ok(out)
    }

    /// `out[i * n_profiled + j] == profiles[j].sw_score_ranges_from_i8(seq_src(seqs[i]))`
    /// (profile_set.rs:313-322): score plus 0-based half-open ranges, no traceback matrix.
    pub fn sw_score_ranges_batch(&self, seqs: &[&[u8]]) -> Result<Vec<MaybeAligned<ScoreAndRanges<u32>>>, CudaError> {
        let (concat, offsets) = pack(seqs);
        let pairs = seqs.len() * self.profiled_lens.len();
        let mut score = vec![0u32; pairs];
        let (mut status, mut tier) = (vec![0u8; pairs], vec![0u8; pairs]);
        let (mut rs, mut re, mut qs, mut qe) = (vec![0u32; pairs], vec![0u32; pairs], vec![0u32; pairs], vec![0u32; pairs]);
        self.check(unsafe {
            sys::zoe_cuda_sw_score_ranges_batch(self.ctx, concat.as_ptr(), offsets.as_ptr(), seqs.len() as u64,
                score.as_mut_ptr(), status.as_mut_ptr(), tier.as_mut_ptr(), rs.as_mut_ptr(), re.as_mut_ptr(),
                qs.as_mut_ptr(), qe.as_mut_ptr())
        }, 0, 0)?;
        Ok((0..pairs).map(|k| match status[k] {
            sys::ZOE_CUDA_SOME => MaybeAligned::Some(ScoreAndRanges {
                score: score[k],
                ref_range: rs[k] as usize..re[k] as usize,
                query_range: qs[k] as usize..qe[k] as usize,
            }),
            sys::ZOE_CUDA_OVERFLOWED => MaybeAligned::Overflowed,
            _ => MaybeAligned::Unmapped,
        }).collect())
    }

    /// Which integer types may answer: `(8, 32, false)` = `sw_*_from_i8` (default), `(16, 32, false)` =
    /// `..._from_i16`, `(32, 32, false)` = `..._from_i32` (profile_set.rs:71-179); `first == last` = a standalone
    /// `StripedProfile<T, N, S>`; `unsigned` = the u8/u16/u32 profiles over `to_biased_matrix()`.
    pub fn set_width_policy(&self, first_bits: i32, last_bits: i32, unsigned: bool) -> Result<(), CudaError> {
        self.check(unsafe { sys::zoe_cuda_set_width_policy(self.ctx, first_bits, last_bits, unsigned as i32) }, 0, 0)
    }

    /// Tuning of the align pipeline (never changes results): 0 auto, 1 full-matrix flags, 2 checkpointed window.
    pub fn set_align_options(&self, mode: i32, checkpoint_log2: i32, slack: i32) -> Result<(), CudaError> {
        self.check(unsafe { sys::zoe_cuda_set_align_options(self.ctx, mode, checkpoint_log2, slack) }, 0, 0)
    }

    /// Upper bound on the device scratch of one align / ranges / 3-pass call (0 = automatic): larger batches are
    /// processed in chunks of streamed sequences.  Never changes results.
    pub fn set_memory_budget(&self, scratch_bytes: u64) -> Result<(), CudaError> {
        self.check(unsafe { sys::zoe_cuda_set_memory_budget(self.ctx, scratch_bytes) }, 0, 0)
    }

    /// The SeqSrc this context's batches correspond to (alignment/mod.rs:157-162).
    pub fn seq_src<'a>(&self, seq: &'a [u8]) -> SeqSrc<&'a [u8]> {
        if self.profiled_is_query { SeqSrc::Reference(seq) } else { SeqSrc::Query(seq) }
    }

    fn check(&self, rc: i32, gap_open: i8, gap_extend: i8) -> Result<(), CudaError> {
        match rc {
            0 => Ok(()),
            sys::ZOE_CUDA_E_EMPTY_SEQUENCE => Err(CudaError::Profile(ProfileError::EmptySequence)),
            sys::ZOE_CUDA_E_GAP_OPEN_RANGE => Err(CudaError::Profile(ProfileError::GapOpenOutOfRange { gap_open })),
            sys::ZOE_CUDA_E_GAP_EXTEND_RANGE => Err(CudaError::Profile(ProfileError::GapExtendOutOfRange { gap_extend })),
            sys::ZOE_CUDA_E_BAD_GAP_WEIGHTS => Err(CudaError::Profile(ProfileError::BadGapWeights { gap_open, gap_extend })),
            code => {
                let message = unsafe { CStr::from_ptr(sys::zoe_cuda_last_error(self.ctx)) }.to_string_lossy().into_owned();
                Err(CudaError::Cuda { code, message })
            }
        }
    }
}

impl Drop for CudaProfiles {
    fn drop(&mut self) {
        unsafe { sys::zoe_cuda_destroy(self.ctx) }
    }
}

/// `out[i] == zoe::alignment::sneaky_snake(references[i], queries[i], threshold)`
/// (src/alignment/sneaky_snake.rs:78-131) on the given devices.
pub fn sneaky_snake_batch(references: &[&[u8]], queries: &[&[u8]], threshold: f32, devices: &[i32]) -> Result<Vec<Option<bool>>, CudaError> {
    assert_eq!(references.len(), queries.len(), "references and queries must pair up");
    let mut ctx = std::ptr::null_mut();
    let rc = unsafe { sys::zoe_cuda_create(&mut ctx, devices.as_ptr(), devices.len() as i32) };
    if rc != 0 {
        return Err(CudaError::Cuda { code: rc, message: "no usable CUDA device".into() });
    }
    let (rb, ro) = pack(references);
    let (qb, qo) = pack(queries);
    let mut out = vec![0u8; references.len().max(1)];
    let rc = unsafe {
        sys::zoe_cuda_sneaky_snake_batch(ctx, rb.as_ptr(), ro.as_ptr(), qb.as_ptr(), qo.as_ptr(), references.len() as u64,
                                         threshold, out.as_mut_ptr())
    };
    let message = if rc != 0 { unsafe { CStr::from_ptr(sys::zoe_cuda_last_error(ctx)) }.to_string_lossy().into_owned() } else { String::new() };
    unsafe { sys::zoe_cuda_destroy(ctx) };
    if rc != 0 {
        return Err(CudaError::Cuda { code: rc, message });
    }
    Ok(out[..references.len()].iter().map(|&v| match v { 0 => Some(false), 1 => Some(true), _ => None }).collect())
}

fn maybe(status: u8, score: u32) -> MaybeAligned<u32> {
    match status {
        sys::ZOE_CUDA_SOME => MaybeAligned::Some(score),
        sys::ZOE_CUDA_OVERFLOWED => MaybeAligned::Overflowed,
        _ => MaybeAligned::Unmapped,
    }
}

fn pack(seqs: &[&[u8]]) -> (Vec<u8>, Vec<u64>) {
    let mut concat = Vec::with_capacity(seqs.iter().map(|s| s.len()).sum::<usize>().max(1));
    let mut offsets = Vec::with_capacity(seqs.len() + 1);
    offsets.push(0u64);
    for s in seqs {
        concat.extend_from_slice(s);
        offsets.push(concat.len() as u64);
    }
    if concat.is_empty() {
        concat.push(0);
    }
    (concat, offsets)
}
