//! cuda_decoder — `flowgger::decoder::Decoder` implementations backed by the B200 batched parser.
//!
//! Mirrors flowgger_b200/csrc/host/flowgger.{hpp,cpp} (the C++ twin that the test-suite drives, because the
//! build environment has no Rust toolchain).  Reference interfaces implemented here:
//!   * `Decoder::decode(&self, line: &str) -> Result<Record, &'static str>`  (src/flowgger/decoder/mod.rs:44-46)
//!   * `CloneBoxedDecoder` via `#[derive(Clone)]`                            (decoder/mod.rs:23-42)
//!   * `Splitter<T>::run`                                                    (src/flowgger/splitter/mod.rs:18-26)
//! All parsing happens on the GPU — including line framing, the UTF-8 check and the unescape of RFC5424 SD values; this
//! crate hands raw blocks to `fg_split_decode` (or packed lines to `fg_decode_batch`) and materialises Records, or, for
//! the rfc5424 -> gelf pair, forwards the records the device already encoded (`fg_decode_encode_gelf`).
#![allow(non_camel_case_types, non_upper_case_globals, dead_code)]

use flowgger::flowgger::config::Config;
use flowgger::flowgger::decoder::Decoder;
use flowgger::flowgger::encoder::Encoder;
use flowgger::flowgger::record::{Record, SDValue, StructuredData};
use flowgger::flowgger::splitter::Splitter;
use std::ffi::{CStr, CString};
use std::io::{stderr, BufRead, BufReader, ErrorKind, Read, Write};
use std::os::raw::c_char;
use std::ptr;
use std::sync::mpsc::SyncSender;
use std::sync::{Arc, Mutex};

mod ffi {
    include!(concat!(env!("OUT_DIR"), "/ffi.rs"));
}
use ffi::*;

/// One GPU context of a fixed format (`fg_ctx`); shared by the clones handed to input threads.
struct Ctx {
    raw: *mut fg_ctx,
    fmt: fg_format,
    suffix: [Option<String>; 5],
}
unsafe impl Send for Ctx {}
impl Drop for Ctx {
    fn drop(&mut self) {
        unsafe { fg_destroy(self.raw) }
    }
}

#[derive(Clone)]
pub struct CudaDecoder {
    ctx: Arc<Mutex<Ctx>>,
}

fn ltsv_type(t: &str) -> Option<i32> {
    match t.to_lowercase().as_ref() {
        "string" => Some(0),
        "bool" => Some(1),
        "f64" => Some(2),
        "i64" => Some(3),
        "u64" => Some(4),
        _ => None,
    }
}

impl CudaDecoder {
    /// `RFC5424Decoder::new(&Config)` / `LTSVDecoder::new` / `GelfDecoder::new` / `RFC3164Decoder::new` replacement,
    /// selected in `flowgger::start` (src/flowgger/mod.rs:413-422) by `input.format`.
    /// RFC3164 (`fg_format_FG_FMT_RFC3164`): the context follows the UTC clock for `now_utc().year()`
    /// (rfc3164_decoder.rs:175, `rfc3164_year = 0`) and reads the zone database behind `timezones::get_by_name` (:196) from
    /// the TZif files under `input.cuda_tzdir` / `$TZDIR` / /usr/share/zoneinfo.
    pub fn new(config: &Config, fmt: fg_format) -> CudaDecoder {
        let mut names: Vec<CString> = Vec::new();
        let mut types: Vec<i32> = Vec::new();
        let mut suffix: [Option<String>; 5] = Default::default();
        let has_schema = config.lookup("input.ltsv_schema").is_some();
        if let Some(pairs) = config.lookup("input.ltsv_schema") {
            // same panics as ltsv_decoder.rs:31-44
            for (name, sdtype) in pairs.as_table().expect("input.ltsv_schema must be a list of key/type pairs") {
                let t = sdtype.as_str().expect("input.ltsv_schema types must be strings");
                let t = ltsv_type(t).unwrap_or_else(|| panic!("Unsupported type in input.ltsv_schema for name [{}]", name));
                names.push(CString::new(name.as_str()).unwrap());
                types.push(t);
            }
        }
        if let Some(pairs) = config.lookup("input.ltsv_suffixes") {
            for (sdtype, sfx) in pairs.as_table().expect("input.ltsv_suffixes must be a list of type/suffixes pairs") {
                let sfx = sfx.as_str().expect("input.ltsv_suffixes suffixes must be strings").to_owned();
                match ltsv_type(sdtype) {
                    Some(0) => panic!("Strings cannot be suffixed"),
                    Some(t) => suffix[t as usize] = Some(sfx),
                    None => panic!("Unsupported type in input.ltsv_suffixes for type [{}]", sdtype),
                }
            }
        }
        let name_ptrs: Vec<*const c_char> = names.iter().map(|s| s.as_ptr()).collect();
        let sfx_c: Vec<Option<CString>> = suffix.iter().map(|s| s.as_ref().map(|x| CString::new(x.as_str()).unwrap())).collect();
        let mut cfg: fg_config = unsafe { std::mem::zeroed() };
        cfg.device = config.lookup("input.cuda_device").and_then(|v| v.as_integer()).unwrap_or(0) as i32;
        cfg.max_batch_bytes = config.lookup("input.cuda_max_batch_bytes").and_then(|v| v.as_integer()).unwrap_or(0);
        cfg.max_batch_lines = config.lookup("input.cuda_max_batch_lines").and_then(|v| v.as_integer()).unwrap_or(0) as i32;
        let tzdir: Option<CString> = config.lookup("input.cuda_tzdir").and_then(|v| v.as_str()).map(|s| CString::new(s).unwrap());
        cfg.rfc3164_year = 0;
        cfg.tzdir = tzdir.as_ref().map_or(ptr::null(), |s| s.as_ptr());
        cfg.ltsv_has_schema = has_schema as i32;
        cfg.ltsv_schema_len = names.len() as i32;
        cfg.ltsv_schema_names = name_ptrs.as_ptr();
        cfg.ltsv_schema_types = types.as_ptr();
        for t in 1..5 {
            cfg.ltsv_suffix[t] = sfx_c[t].as_ref().map_or(ptr::null(), |s| s.as_ptr());
        }
        let mut raw: *mut fg_ctx = ptr::null_mut();
        let rc = unsafe { fg_create(&cfg, &mut raw) };
        if rc != 0 {
            // There is no CPU fallback: without the library + a GPU the decoder cannot exist.
            panic!("flowgger_cuda: fg_create failed ({})", rc);
        }
        CudaDecoder { ctx: Arc::new(Mutex::new(Ctx { raw, fmt, suffix })) }
    }

    /// Decode `n` lines packed as bytes + offsets; calls `f(i, result)` in input order.
    pub fn decode_batch<F: FnMut(usize, Result<Record, &'static str>, &[String])>(&self, bytes: &[u8], offsets: &[i32], mut f: F) {
        let ctx = self.ctx.lock().unwrap();
        let n = offsets.len() - 1;
        let mut out: fg_batch_out = unsafe { std::mem::zeroed() };
        let rc = unsafe { fg_decode_batch(ctx.raw, ctx.fmt, bytes.as_ptr(), offsets.as_ptr(), n as i32, &mut out) };
        if rc != 0 {
            let e = unsafe { CStr::from_ptr(fg_last_error(ctx.raw)) }.to_string_lossy().into_owned();
            panic!("fg_decode_batch: {}", e);
        }
        for i in 0..n {
            let mut side = Vec::new();
            let r = materialize(&ctx, &out, bytes, offsets, i, &mut side);
            f(i, r, &side);
        }
    }
}

impl CudaDecoder {
    /// Framing + UTF-8 validation + decode of a raw stream on the device (`fg_split_decode`): `f(i, line, result)` in
    /// stream order; a line that is not UTF-8 arrives as `Err("Invalid UTF-8 input")` (line_splitter.rs:22-25).
    pub fn split_decode<F: FnMut(usize, &[u8], Result<Record, &'static str>, &[String])>(&self, stream: &[u8], mut f: F) {
        let ctx = self.ctx.lock().unwrap();
        let mut out: fg_batch_out = unsafe { std::mem::zeroed() };
        let rc = unsafe { fg_split_decode(ctx.raw, ctx.fmt, stream.as_ptr(), stream.len() as i64, &mut out) };
        if rc != 0 {
            let e = unsafe { CStr::from_ptr(fg_last_error(ctx.raw)) }.to_string_lossy().into_owned();
            panic!("fg_split_decode: {}", e);
        }
        let offs = unsafe { std::slice::from_raw_parts(out.line_offsets, out.n as usize + 1) };
        for i in 0..out.n as usize {
            // BufRead::lines: drop the '\n' and one '\r' before it
            let (lo, mut hi) = (offs[i] as usize, offs[i + 1] as usize);
            if hi > lo && stream[hi - 1] == b'\n' { hi -= 1; if hi > lo && stream[hi - 1] == b'\r' { hi -= 1; } }
            let ext = [lo as i32, hi as i32];
            let mut side = Vec::new();
            // `materialize` reads the extent of line i from offsets[i..i+2]: hand it the stripped extent
            let r = materialize_ext(&ctx, &out, stream, ext[0], ext[1], i, &mut side);
            f(i, &stream[lo..hi], r, &side);
        }
    }
}

fn span<'a>(bytes: &'a [u8], s: fg_span) -> &'a str {
    // spans delimit whole UTF-8 sequences of a line that was validated before the call
    unsafe { std::str::from_utf8_unchecked(&bytes[s.off as usize..(s.off + s.len) as usize]) }
}

/// JSON string body already validated on the device -> String; `nl_retry` = gelf_decoder.rs:44-46 semantics.
fn json_unescape(v: &str, nl_retry: bool) -> String {
    let b = v.as_bytes();
    let mut out = String::with_capacity(b.len());
    let hex = |s: &[u8]| s.iter().fold(0u32, |n, &c| n * 16 + (c as char).to_digit(16).unwrap());
    let mut i = 0;
    let mut raw_from = 0;
    while i < b.len() {
        if b[i] != b'\\' { i += 1; continue; }
        out.push_str(&v[raw_from..i]);
        let e = b[i + 1];
        i += 2;
        match e {
            b'"' => out.push('"'), b'\\' => out.push('\\'), b'/' => out.push('/'),
            b'b' => out.push('\x08'), b'f' => out.push('\x0c'), b'n' => out.push('\n'),
            b'r' => out.push('\r'), b't' => out.push('\t'),
            b'u' => {
                let mut n = hex(&b[i..i + 4]);
                i += 4;
                if (0xD800..=0xDBFF).contains(&n) {
                    let n2 = hex(&b[i + 2..i + 6]);
                    i += 6;
                    n = (((n - 0xD800) << 10) | (n2 - 0xDC00)) + 0x1_0000;
                }
                out.push(std::char::from_u32(n).unwrap());
            }
            b'\n' if nl_retry => out.push_str("\\n"),
            _ => {}
        }
        raw_from = i;
    }
    out.push_str(&v[raw_from..]);
    out
}

fn materialize(ctx: &Ctx, out: &fg_batch_out, bytes: &[u8], offsets: &[i32], i: usize, side: &mut Vec<String>) -> Result<Record, &'static str> {
    materialize_ext(ctx, out, bytes, offsets[i], offsets[i + 1], i, side)
}

/// RFC5424: compact 32-byte row (`fg_row5424`) + 8-byte entries; escaped SD values were unescaped ON THE DEVICE
/// (rfc5424_decoder.rs:105-125) and live in `out.arena`.  Rows flagged FG_FLAG_WIDE carry absolute spans instead.
unsafe fn materialize_5424(ctx: &Ctx, out: &fg_batch_out, bytes: &[u8], line_lo: i32, i: usize) -> Result<Record, &'static str> {
    let row = &*out.rows5424.add(i);
    let status = row.meta & 0xFF;
    if status != 0 {
        let s = CStr::from_ptr(fg_error_string(ctx.fmt, status));
        return Err(std::str::from_utf8_unchecked(std::slice::from_raw_parts(s.as_ptr() as *const u8, s.to_bytes().len())));
    }
    let (fac, sev, flags) = ((row.meta >> 8) & 0xFF, (row.meta >> 16) & 0xFF, row.meta >> 24);
    let arena = |off: u32, len: usize| std::str::from_utf8_unchecked(std::slice::from_raw_parts(out.arena.add(off as usize), len)).to_owned();
    if flags & FG_FLAG_WIDE != 0 {
        let w = &*out.wide_rows.add(row.sd_first as usize);
        let mut sd: Vec<StructuredData> = Vec::new();
        for e in w.sd.off..w.sd.off + w.sd.len {
            let e = e as usize;
            let em = *out.entry_meta.add(e) as u32;
            let nm = *out.entry_name.add(e);
            if em & FG_EM_TAG_MASK == fg_tag_FG_TAG_SD_HEADER as u32 {
                sd.push(StructuredData::new(Some(span(bytes, nm))));
                continue;
            }
            let val = *out.entry_val.add(e);
            let v = if em & FG_EM_ARENA != 0 { arena(val as u32, (val >> 32) as usize) }
                    else { span(bytes, fg_span { off: (val & 0xFFFF_FFFF) as i32, len: (val >> 32) as i32 }).to_owned() };
            sd.last_mut().unwrap().pairs.push((format!("_{}", span(bytes, nm)), SDValue::String(v)));
        }
        return Ok(Record {
            ts: w.ts, hostname: span(bytes, w.hostname).to_owned(), facility: Some(fac as u8), severity: Some(sev as u8),
            appname: Some(span(bytes, w.appname).to_owned()), procid: Some(span(bytes, w.procid).to_owned()),
            msgid: Some(span(bytes, w.msgid).to_owned()),
            msg: if w.msg.off >= 0 { Some(span(bytes, w.msg).to_owned()) } else { None },
            full_msg: Some(span(bytes, w.full_msg).to_owned()),
            sd: if sd.is_empty() { None } else { Some(sd) },
        });
    }
    let line = &bytes[line_lo as usize..];
    let rel = |a: usize, b: usize| std::str::from_utf8_unchecked(&line[a..b]).to_owned();
    let sp: Vec<usize> = row.sp.iter().map(|&x| x as usize).collect();
    let (mo, ml) = (row.msg_off as usize, row.msg_len as usize);
    let mut sd: Vec<StructuredData> = Vec::new();
    let mut e = row.sd_first as usize;
    let end = e + row.sd_count as usize;
    while e < end {
        let v = *out.entries8.add(e);
        let (a, b, c) = ((v & 0xFFFF) as usize, ((v >> 16) & 0xFFFF) as usize, ((v >> 32) & 0xFFFF) as usize);
        if v & FG_E8_HEADER != 0 {
            sd.push(StructuredData::new(Some(&rel(a, b))));
        } else {
            let value = if v & FG_E8_ARENA != 0 {
                let off = (((v >> 32) & 0x3FFF_FFFF) as u32) << 1;   // record = [u16 length][bytes]
                let len = u16::from_le_bytes([*out.arena.add(off as usize), *out.arena.add(off as usize + 1)]) as usize;
                arena(off + 2, len)
            } else {
                rel(b + 2, c)
            };
            sd.last_mut().unwrap().pairs.push((format!("_{}", rel(a, b)), SDValue::String(value)));   // :221
        }
        e += 1;
    }
    Ok(Record {
        ts: row.ts, hostname: rel(sp[0] + 1, sp[1]), facility: Some(fac as u8), severity: Some(sev as u8),
        appname: Some(rel(sp[1] + 1, sp[2])), procid: Some(rel(sp[2] + 1, sp[3])), msgid: Some(rel(sp[3] + 1, sp[4])),
        msg: if ml > 0 { Some(rel(mo, mo + ml)) } else { None },
        full_msg: Some(rel(0, mo + ml)),
        sd: if sd.is_empty() { None } else { Some(sd) },
    })
}

fn materialize_ext(ctx: &Ctx, out: &fg_batch_out, bytes: &[u8], line_lo: i32, line_hi: i32, i: usize, side: &mut Vec<String>) -> Result<Record, &'static str> {
    unsafe {
        if !out.rows5424.is_null() {
            return materialize_5424(ctx, out, bytes, line_lo, i);
        }
        let meta = *out.meta.add(i);
        let status = meta & 0xFF;
        let flags = (meta >> 24) & 0xFF;
        if flags & FG_FLAG_MISSING_VALUE != 0 {
            // println! at ltsv_decoder.rs:99 for every part without ':' that the decode loop reached
            let (lo, hi) = (line_lo as usize, line_hi as usize);
            let stop = if status != 0 { (*out.full_msg.add(i)).off as usize } else { hi + 1 };
            let mut a = lo;
            for part in span(bytes, fg_span { off: lo as i32, len: (hi - lo) as i32 }).split('\t') {
                if a >= stop { break; }
                if !part.contains(':') { side.push(format!("Missing value for name '{}'", part)); }
                a += part.len() + 1;
            }
        }
        if status != 0 {
            let s = CStr::from_ptr(fg_error_string(ctx.fmt, status));
            return Err(std::str::from_utf8_unchecked(std::slice::from_raw_parts(s.as_ptr() as *const u8, s.to_bytes().len())));
        }
        let nl = flags & FG_FLAG_NL_RETRY != 0;
        let opt = |col: *const fg_span| if col.is_null() || (*col.add(i)).off < 0 { None } else { Some(span(bytes, *col.add(i)).to_owned()) };
        let esc = |s: Option<String>, bit: u32| s.map(|x| if flags & bit != 0 { json_unescape(&x, nl) } else { x });
        let fac = (meta >> 8) & 0xFF;
        let sev = (meta >> 16) & 0xFF;
        let sd_span = *out.sd.add(i);
        let mut sd: Vec<StructuredData> = Vec::new();
        if sd_span.len > 0 {
            sd.push(StructuredData::new(None));  // LTSV / GELF: one element without sd_id
            for e in sd_span.off..sd_span.off + sd_span.len {
                let e = e as usize;
                let em = *out.entry_meta.add(e) as u32;
                let tag = em & FG_EM_TAG_MASK;
                let nm = *out.entry_name.add(e);
                if tag == fg_tag_FG_TAG_SD_HEADER as u32 {
                    sd.push(StructuredData::new(if nm.off >= 0 { Some(span(bytes, nm)) } else { None }));
                    continue;
                }
                let mut name = String::new();
                if em & FG_EM_NO_PREFIX == 0 { name.push('_'); }
                if em & FG_EM_NAME_ESC != 0 { name.push_str(&json_unescape(span(bytes, nm), nl)); } else { name.push_str(span(bytes, nm)); }
                if em & FG_EM_SUFFIX != 0 { if let Some(ref s) = ctx.suffix[tag as usize] { name.push_str(s); } }
                let val = *out.entry_val.add(e);
                let v = match tag {
                    0 => {
                        let raw = span(bytes, fg_span { off: (val & 0xFFFF_FFFF) as i32, len: (val >> 32) as i32 });
                        SDValue::String(if em & FG_EM_UNESCAPE == 0 { raw.to_owned() } else { json_unescape(raw, nl) })
                    }
                    1 => SDValue::Bool(val != 0),
                    2 => SDValue::F64(f64::from_bits(val)),
                    3 => SDValue::I64(val as i64),
                    4 => SDValue::U64(val),
                    _ => SDValue::Null,
                };
                sd.last_mut().unwrap().pairs.push((name, v));
            }
        }
        let ts = if flags & FG_FLAG_TS_MISSING != 0 {
            flowgger::flowgger::utils::PreciseTimestamp::now().as_f64() // gelf_decoder.rs:109
        } else {
            *out.ts.add(i)
        };
        Ok(Record {
            ts,
            hostname: esc(opt(out.hostname), FG_FLAG_HOST_ESC).unwrap_or_default(),
            facility: if fac != 0xFF { Some(fac as u8) } else { None },
            severity: if sev != 0xFF { Some(sev as u8) } else { None },
            appname: opt(out.appname),
            procid: opt(out.procid),
            msgid: opt(out.msgid),
            msg: if flags & FG_FLAG_MSG_ARENA != 0 {
                // RFC3164: the message tokens re-joined by single spaces on the device (rfc3164_decoder.rs:67)
                let m = *out.msg.add(i);
                Some(std::str::from_utf8_unchecked(std::slice::from_raw_parts(out.arena.add(m.off as usize), m.len as usize)).to_owned())
            } else {
                esc(opt(out.msg), FG_FLAG_MSG_ESC)
            },
            full_msg: esc(opt(out.full_msg), FG_FLAG_FULL_ESC),
            sd: if sd.is_empty() { None } else { Some(sd) },
        })
    }
}

impl Decoder for CudaDecoder {
    /// Drop-in single-line form: a batch of one through the same kernels.
    fn decode(&self, line: &str) -> Result<Record, &'static str> {
        let offsets = [0i32, line.len() as i32];
        let mut res = Err("unreachable");
        self.decode_batch(line.as_bytes(), &offsets, |_, r, side| {
            for s in side { println!("{}", s); }
            res = r;
        });
        res
    }
}

/// Batched twin of `LineSplitter` (src/flowgger/splitter/line_splitter.rs:10-54): inserted between `input` and
/// `decoder`.  It reads RAW BLOCKS (no per-line `String`), cuts each block after its last '\n' and hands the block to
/// `fg_split_decode`: line framing (`BufRead::lines`: "\n", one "\r"), the UTF-8 check of `String` and the decode all run
/// on the device; stderr text and record order are those of the reference.
pub struct BatchingLineSplitter {
    pub gpu: CudaDecoder,
    pub max_bytes: usize,
}

impl BatchingLineSplitter {
    fn flush<F: FnMut(Vec<u8>)>(&self, block: &[u8], encoder: &Box<dyn Encoder>, send: &mut F) {
        self.gpu.split_decode(block, |_, line, r, side| {
            for s in side { println!("{}", s); }
            match r.and_then(|rec| encoder.encode(rec)) {
                Ok(bytes) => send(bytes),
                Err("Invalid UTF-8 input") => { let _ = writeln!(stderr(), "Invalid UTF-8 input"); }   // line_splitter.rs:22-25
                Err(e) => { let _ = writeln!(stderr(), "{}: [{}]", e, String::from_utf8_lossy(line).trim()); }  // :37-39
            }
        });
    }
}

impl<T: Read> Splitter<T> for BatchingLineSplitter {
    fn run(&self, mut buf_reader: BufReader<T>, tx: SyncSender<Vec<u8>>, _decoder: Box<dyn Decoder>, encoder: Box<dyn Encoder>) {
        let mut block: Vec<u8> = Vec::with_capacity(self.max_bytes);
        let mut send = |bytes: Vec<u8>| tx.send(bytes).unwrap();
        loop {
            let got = match buf_reader.fill_buf() {
                Ok(b) => { let n = b.len(); block.extend_from_slice(b); n }
                Err(e) => match e.kind() {
                    ErrorKind::Interrupted => continue,
                    ErrorKind::WouldBlock => {
                        self.flush(&block, &encoder, &mut send);
                        let _ = writeln!(stderr(), "Client hasn't sent any data for a while - Closing idle connection");
                        return;
                    }
                    _ => { self.flush(&block, &encoder, &mut send); return; }
                },
            };
            buf_reader.consume(got);
            if got == 0 {                       // EOF: an unterminated last line is still a line
                self.flush(&block, &encoder, &mut send);
                return;
            }
            if block.len() >= self.max_bytes {
                // decode every complete line of the block, keep the unterminated tail for the next one
                let cut = block.iter().rposition(|&c| c == b'\n').map_or(0, |p| p + 1);
                if cut > 0 {
                    self.flush(&block[..cut], &encoder, &mut send);
                    block.drain(..cut);
                }
            }
        }
    }
}

/// `output.format = "gelf"` with `input.format = "rfc5424"`: decode AND encode run on the device
/// (`fg_decode_encode_gelf`, replaces Decoder::decode + GelfEncoder::encode of line_splitter.rs:50-52); only the encoded
/// records come back.  Lines are framed on the host here (the fused entry point takes offsets).
pub struct FusedGelfLineSplitter {
    pub gpu: CudaDecoder,
    pub extra: Vec<(String, String)>,   // output.gelf_extra (gelf_encoder.rs:29-48)
    pub max_lines: usize,
    pub max_bytes: usize,
}

impl<T: Read> Splitter<T> for FusedGelfLineSplitter {
    fn run(&self, buf_reader: BufReader<T>, tx: SyncSender<Vec<u8>>, _decoder: Box<dyn Decoder>, _encoder: Box<dyn Encoder>) {
        let ctx = self.gpu.ctx.lock().unwrap();
        let keys: Vec<CString> = self.extra.iter().map(|(k, _)| CString::new(k.as_str()).unwrap()).collect();
        let vals: Vec<CString> = self.extra.iter().map(|(_, v)| CString::new(v.as_str()).unwrap()).collect();
        let kp: Vec<*const c_char> = keys.iter().map(|s| s.as_ptr()).collect();
        let vp: Vec<*const c_char> = vals.iter().map(|s| s.as_ptr()).collect();
        assert_eq!(unsafe { fg_set_gelf_extra(ctx.raw, kp.len() as i32, kp.as_ptr(), vp.as_ptr()) }, 0);
        let mut arena: Vec<u8> = Vec::with_capacity(self.max_bytes);
        let mut offsets: Vec<i32> = vec![0];
        let flush = |arena: &mut Vec<u8>, offsets: &mut Vec<i32>| {
            if offsets.len() > 1 {
                let mut out: fg_encoded_out = unsafe { std::mem::zeroed() };
                let rc = unsafe { fg_decode_encode_gelf(ctx.raw, ctx.fmt, arena.as_ptr(), offsets.as_ptr(), offsets.len() as i32 - 1, &mut out) };
                assert_eq!(rc, 0);
                for i in 0..out.n as usize {
                    let (st, lo, hi) = unsafe { (*out.status.add(i), *out.offsets.add(i) as usize, *out.offsets.add(i + 1) as usize) };
                    if st == 0 {
                        tx.send(unsafe { std::slice::from_raw_parts(out.bytes.add(lo), hi - lo) }.to_vec()).unwrap();
                    } else {
                        let e = unsafe { CStr::from_ptr(fg_error_string(ctx.fmt, st as u32)) }.to_string_lossy();
                        let line = String::from_utf8_lossy(&arena[offsets[i] as usize..offsets[i + 1] as usize]);
                        let _ = writeln!(stderr(), "{}: [{}]", e, line.trim());
                    }
                }
            }
            arena.clear();
            offsets.clear();
            offsets.push(0);
        };
        for line in buf_reader.lines() {
            match line {
                Ok(line) => {
                    if arena.len() + line.len() > self.max_bytes || offsets.len() > self.max_lines { flush(&mut arena, &mut offsets); }
                    arena.extend_from_slice(line.as_bytes());
                    offsets.push(arena.len() as i32);
                }
                Err(e) => match e.kind() {
                    ErrorKind::Interrupted => continue,
                    ErrorKind::InvalidInput | ErrorKind::InvalidData => { flush(&mut arena, &mut offsets); let _ = writeln!(stderr(), "Invalid UTF-8 input"); }
                    _ => break,
                },
            }
        }
        flush(&mut arena, &mut offsets);
    }
}
