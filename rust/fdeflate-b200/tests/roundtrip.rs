//! The reference's own unit tests for this path (src/decompress.rs:1235-1325, src/compress/ultrafast.rs:195-224),
//! run against the device-backed API.  Needs a CUDA device.
use fdeflate_b200::*;

fn data() -> Vec<u8> {
    let mut s = 12345u32;
    (0..200_000).map(|_| {
        s = s.wrapping_mul(1664525).wrapping_add(1013904223);
        if (s >> 24) < 120 { 0 } else { ((s >> 16) % 7) as u8 }
    }).collect()
}

#[test]
fn ultrafast_roundtrip_through_miniz() {
    let d = data();
    let z = compress_to_vec_ultra_fast(&d);
    assert_eq!(miniz_oxide::inflate::decompress_to_vec_zlib(&z).unwrap(), d);
    assert_eq!(decompress_to_vec(&z).unwrap(), d);
}

#[test]
fn multi_call_matches_single_call_on_chunk_edges() {
    let d: Vec<u8> = data().into_iter().map(|b| b | 1).collect(); // no zero runs: cuts on 8-byte edges are invisible
    let mut c = UltraFastCompressor::new(Vec::new()).unwrap();
    c.write_data(&d[..1024]).unwrap();
    c.write_data(&d[1024..]).unwrap();
    assert_eq!(c.finish().unwrap(), compress_to_vec_ultra_fast(&d));
}

#[test]
fn wrong_checksum_and_ignore() {
    let mut z = compress_to_vec_ultra_fast(b"Hello world!");
    *z.last_mut().unwrap() ^= 1;
    assert_eq!(decompress_to_vec(&z), Err(DecompressionError::WrongChecksum));
    let mut dec = Decompressor::new();
    dec.ignore_adler32();
    let mut out = vec![0u8; 64];
    let (_, n) = dec.read(&z, &mut out, 0).unwrap();
    assert!(dec.is_done());
    assert_eq!(&out[..n], b"Hello world!");
}

#[test]
fn bytewise_equals_whole() {
    let d = data();
    let z = compress_to_vec_ultra_fast(&d[..20_000]);
    let mut dec = Decompressor::new();
    let mut out = vec![0u8; 30_000];
    let mut pos = 0;
    for b in z.chunks(1) {
        let (_, n) = dec.read(b, &mut out, pos).unwrap();
        pos += n;
    }
    assert!(dec.is_done());
    assert_eq!(&out[..pos], &d[..20_000]);
}

#[test]
fn bounded() {
    let d = data();
    let z = compress_to_vec_ultra_fast(&d);
    assert!(decompress_to_vec_bounded(&z, d.len()).is_ok());
    match decompress_to_vec_bounded(&z, d.len() - 1) {
        Err(BoundedDecompressionError::OutputTooLarge { partial_output }) => assert_eq!(partial_output, &d[..d.len() - 1]),
        _ => panic!("expected OutputTooLarge"),
    }
}
