"""Host-side mirror of the reference's CLI pieces the golden files are written
against (bin/dn argument parsing and printers, lib/attr-parser.js,
lib/path-enum.js): test infrastructure, kept out of the product package."""
