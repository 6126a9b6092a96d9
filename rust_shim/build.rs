fn main() {
    let dir = std::env::var("PHYSIM_B200_LIB_DIR").unwrap_or_else(|_| "../physim_b200".into());
    println!("cargo:rustc-link-search=native={dir}");
    println!("cargo:rustc-link-lib=dylib=physim_b200");
}
