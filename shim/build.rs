// Links libb2rsa.so (built by `make -C halo2-rsa_b200`): LIBB2RSA_DIR points at halo2-rsa_b200/lib.
fn main() {
    let dir = std::env::var("LIBB2RSA_DIR").unwrap_or_else(|_| "../halo2-rsa_b200/lib".to_string());
    println!("cargo:rustc-link-search=native={}", dir);
    println!("cargo:rustc-link-lib=dylib=b2rsa");
    println!("cargo:rustc-link-arg=-Wl,-rpath,{}", dir);
    println!("cargo:rerun-if-env-changed=LIBB2RSA_DIR");
}
