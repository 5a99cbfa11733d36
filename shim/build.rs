// links libphnsw.so; PHNSW_LIB_DIR points at parallel_hnsw_b200/ (where _build.py leaves it)
fn main() {
    let dir = std::env::var("PHNSW_LIB_DIR").unwrap_or_else(|_| "../parallel_hnsw_b200".to_string());
    println!("cargo:rustc-link-search=native={dir}");
    println!("cargo:rustc-link-lib=dylib=phnsw");
    println!("cargo:rerun-if-env-changed=PHNSW_LIB_DIR");
}
