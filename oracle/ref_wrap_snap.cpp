// TEST INFRASTRUCTURE ONLY -- wrappers around the reference's snaplegacy library (filled in below).
