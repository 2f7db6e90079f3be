{-# LANGUAGE ForeignFunctionInterface #-}
-- | SkaGridFFI -- the reference-side binding of libskagrid.so (include/skagrid.h).
--
-- This is the module a maintainer of sakehl/SKA-SDP-Accelerate-gridding adds next to src/Hdf5.hs: it follows
-- that file's FFI idiom (src/Hdf5.hs:30-67: `foreign import ccall`, CString/Ptr arguments, buffers allocated
-- on the Haskell side) and re-exports the names of src/Gridding.hs with host arrays instead of `Acc` terms.
-- With it, `ImageDataset.aw_gridding` (src/ImageDataset.hs:29-83) stops building an Accelerate AST for the
-- hot path and calls `awGridding` below; `Runners` (src/Gridding.hs:28) is no longer needed on that path.
--
-- NOT COMPILED in the build environment of this repository (no GHC available there); it is written against
-- the Accelerate / accelerate-io versions pinned in the reference's CPUstack.yaml:12-34.
--
-- `safe` imports are used on purpose: the calls block for milliseconds to seconds (H2D copy, kernels, D2H
-- copy) and an `unsafe` call would stall the RTS capability (SURVEY.md section 8b).
module SkaGridFFI
  ( Ctx, withSkaGrid
  , fracCoords, findClosestV, mirrorUvw, doweight
  , convgrid, convgrid2, convgridAW, convdegrid2, convdegridAW
  , withSkaGrids, MultiMode(..), convgrid2Multi, convdegrid2Multi
  , makeGridHermitian, ifft, gridToImage, awImaging, awGridding
  , gridSide, wKernel
  ) where

import Control.Exception (bracket)
import Control.Monad (when)
import Data.Array.Accelerate (Array, DIM1, DIM2, DIM3, DIM5, Vector, Z(..), (:.)(..), arrayShape, arraySize)
import Data.Array.Accelerate.Data.Complex (Complex)
import Data.Array.Accelerate.IO.Foreign.ForeignPtr (fromForeignPtrs, toForeignPtrs)
import Data.Int (Int64)
import Foreign.C.String (CString, peekCString)
import Foreign.C.Types (CDouble(..), CInt(..))
import Foreign.ForeignPtr (ForeignPtr, castForeignPtr, mallocForeignPtrArray, withForeignPtr)
import Foreign.Marshal.Alloc (alloca)
import Foreign.Marshal.Array (withArrayLen)
import Foreign.Ptr (Ptr, nullPtr)
import Foreign.Storable (peek)

data CtxS
type Ctx = Ptr CtxS

foreign import ccall safe "skagrid_create"       c_create      :: CInt -> Ptr Ctx -> IO CInt
foreign import ccall safe "skagrid_destroy"      c_destroy     :: Ctx -> IO ()
foreign import ccall safe "skagrid_last_error"   c_last_error  :: Ctx -> IO CString

foreign import ccall safe "skagrid_frac_coords"  c_frac_coords
  :: Ctx -> Int64 -> Int64 -> Int64 -> Int64 -> Ptr Double -> Ptr Double
  -> Ptr Int64 -> Ptr Int64 -> Ptr Int64 -> Ptr Int64 -> CInt -> IO CInt
foreign import ccall safe "skagrid_find_closest" c_find_closest
  :: Ctx -> Int64 -> Ptr Double -> Int64 -> Ptr Double -> Ptr Int64 -> IO CInt
foreign import ccall safe "skagrid_mirror_uvw"   c_mirror_uvw
  :: Ctx -> Int64 -> Ptr Double -> Ptr Double -> Ptr Double -> Ptr Double -> IO CInt
foreign import ccall safe "skagrid_doweight"     c_doweight
  :: Ctx -> CDouble -> Int64 -> Int64 -> Ptr Double -> Ptr Double -> Ptr Double -> IO CInt
foreign import ccall safe "skagrid_convgrid"     c_convgrid
  :: Ctx -> Int64 -> Int64 -> Int64 -> Ptr Double -> Int64 -> Int64 -> Ptr Double
  -> Int64 -> Ptr Double -> Ptr Double -> Ptr Double -> IO CInt
foreign import ccall safe "skagrid_convgrid2"    c_convgrid2
  :: Ctx -> Int64 -> Int64 -> Int64 -> Int64 -> Ptr Double -> Int64 -> Int64 -> Ptr Double
  -> Int64 -> Ptr Double -> Ptr Double -> Ptr Int64 -> Ptr Double -> IO CInt
foreign import ccall safe "skagrid_convgrid_aw"  c_convgrid_aw
  :: Ctx -> Int64 -> Int64 -> Int64 -> Ptr Double -> Int64 -> Ptr Double -> Int64 -> Int64 -> Ptr Double
  -> Int64 -> Ptr Double -> Ptr Double -> Ptr Int64 -> Ptr Int64 -> Ptr Int64 -> Ptr Double -> IO CInt
foreign import ccall safe "skagrid_convdegrid2"  c_convdegrid2
  :: Ctx -> Int64 -> Int64 -> Int64 -> Int64 -> Ptr Double -> Int64 -> Int64 -> Ptr Double
  -> Int64 -> Ptr Double -> Ptr Double -> Ptr Int64 -> Ptr Double -> IO CInt
foreign import ccall safe "skagrid_convdegrid_aw" c_convdegrid_aw
  :: Ctx -> Int64 -> Int64 -> Int64 -> Ptr Double -> Int64 -> Ptr Double -> Int64 -> Int64 -> Ptr Double
  -> Int64 -> Ptr Double -> Ptr Double -> Ptr Int64 -> Ptr Int64 -> Ptr Int64 -> Ptr Double -> IO CInt
-- single-process multi-GPU: an array of contexts, one per device (include/skagrid.h, "multi-GPU, single process")
foreign import ccall safe "skagrid_convgrid2_mgpu_vis"    c_convgrid2_mgpu_vis
  :: Ptr Ctx -> CInt -> Int64 -> Int64 -> Int64 -> Int64 -> Ptr Double -> Int64 -> Int64 -> Ptr Double
  -> Int64 -> Ptr Double -> Ptr Double -> Ptr Int64 -> Ptr Double -> IO CInt
foreign import ccall safe "skagrid_convdegrid2_mgpu_vis"  c_convdegrid2_mgpu_vis
  :: Ptr Ctx -> CInt -> Int64 -> Int64 -> Int64 -> Int64 -> Ptr Double -> Int64 -> Int64 -> Ptr Double
  -> Int64 -> Ptr Double -> Ptr Double -> Ptr Int64 -> Ptr Double -> IO CInt
foreign import ccall safe "skagrid_convgrid2_mgpu_tile"   c_convgrid2_mgpu_tile
  :: Ptr Ctx -> CInt -> Int64 -> Int64 -> Int64 -> Int64 -> Ptr Double -> Int64 -> Int64 -> Ptr Double
  -> Int64 -> Ptr Double -> Ptr Double -> Ptr Int64 -> Ptr Double -> Ptr Int64 -> IO CInt
foreign import ccall safe "skagrid_convdegrid2_mgpu_tile" c_convdegrid2_mgpu_tile
  :: Ptr Ctx -> CInt -> Int64 -> Int64 -> Int64 -> Int64 -> Ptr Double -> Int64 -> Int64 -> Ptr Double
  -> Int64 -> Ptr Double -> Ptr Double -> Ptr Int64 -> Ptr Double -> Ptr Int64 -> IO CInt
foreign import ccall safe "skagrid_make_grid_hermitian" c_hermitian :: Ctx -> Int64 -> Ptr Double -> Ptr Double -> IO CInt
foreign import ccall safe "skagrid_ifft"         c_ifft        :: Ctx -> Int64 -> Ptr Double -> Ptr Double -> IO CInt
foreign import ccall safe "skagrid_grid_to_image" c_grid_to_image
  :: Ctx -> Int64 -> Ptr Double -> Ptr Double -> Ptr Double -> IO CInt
foreign import ccall safe "skagrid_aw_imaging"   c_aw_imaging
  :: Ctx -> CDouble -> Int64 -> Int64 -> Int64 -> Int64 -> Ptr Double -> Ptr Double -> Int64 -> Ptr Double
  -> Int64 -> Ptr Double -> Ptr Double -> Ptr Double -> Ptr Int64 -> Ptr Int64 -> Ptr Double -> Ptr Double -> IO CInt
foreign import ccall safe "skagrid_aw_gridding"  c_aw_gridding
  :: Ctx -> CDouble -> Int64 -> Int64 -> Int64 -> Int64 -> Ptr Double -> Ptr Double -> Int64 -> Ptr Double
  -> Int64 -> Ptr Double -> Ptr Double -> Ptr Double -> Ptr Int64 -> Ptr Int64 -> CDouble -> Ptr Double
  -> Ptr Double -> Ptr Double -> Ptr Double -> IO CInt

foreign import ccall unsafe "skagrid_grid_side"  c_grid_side   :: CDouble -> Int64 -> Int64   -- pure arithmetic: no blocking, `unsafe` is fine
foreign import ccall safe "skagrid_w_kernels_ex" c_w_kernels_ex
  :: Ctx -> CDouble -> Int64 -> Ptr Double -> Int64 -> Int64 -> Int64 -> CInt -> Ptr Double -> CDouble -> CDouble -> Ptr Double -> IO CInt

type F = Double
type Visibility = Complex Double

-- | N = P.round (theta * lam) (src/Gridding.hs:466), the side every imaging entry point of the library allocates
gridSide :: F -> Int -> Int
gridSide theta lam = fromIntegral (c_grid_side (realToFrac theta) (fromIntegral lam))

-- | w_kernel (src/Gridding.hs:610-619) for a list of w, with the KernelOptions fields that move the far-field coordinates
-- (kernel_coordinates, :620-635): (patHorShift, patVerShift) and patTransMat as (t00, t01, t10, t11).
-- -> [nw, qpx, qpx, npixKern, npixKern]; conjugate = True is what w_cache_imaging asks for (:441).
wKernel :: Ctx -> F -> [F] -> (Int, Int, Int) -> (Int, Int) -> Maybe (F, F, F, F) -> Bool -> IO (Array DIM5 Visibility)
wKernel ctx theta ws (qpx, npixFF, npixKern) (dl, dm) tmat conj = do
  let nw = length ws
      total = nw * qpx * qpx * npixKern * npixKern
  out <- newCplx total
  withArrayLen (map realToFrac ws :: [Double]) $ \_ pw -> withForeignPtr out $ \po ->
    let call pt = c_w_kernels_ex ctx (realToFrac theta) (fromIntegral nw) pw (fromIntegral npixFF) (fromIntegral npixKern) (fromIntegral qpx)
                                 (if conj then 1 else 0) pt (fromIntegral dl) (fromIntegral dm) po >>= check ctx "w_kernels_ex"
    in case tmat of
         Nothing -> call nullPtr
         Just (a, b, c, d) -> withArrayLen [a, b, c, d] $ \_ pt -> call pt
  return (fromForeignPtrs (Z :. nw :. qpx :. qpx :. npixKern :. npixKern) (castForeignPtr out))

-- | One context per OS thread / GPU (include/skagrid.h "Threading").
withSkaGrid :: Int -> (Ctx -> IO a) -> IO a
withSkaGrid dev = bracket open c_destroy
  where open = alloca $ \p -> do
          rc <- c_create (fromIntegral dev) p
          when (rc /= 0) $ c_last_error nullPtr >>= peekCString >>= \m -> error ("skagrid_create: " ++ m)
          peek p

check :: Ctx -> String -> CInt -> IO ()
check ctx what rc = when (rc /= 0) $ c_last_error ctx >>= peekCString >>= \m -> error (what ++ ": " ++ m)

-- Complex arrays have the single-buffer (re,im) representation the reference already relies on
-- (src/Hdf5.hs:165-167 unsafeCastDataSet; hdf5/hdf5.cc:14-17 complexDouble{r,i}).
cplxPtr :: Array sh Visibility -> ForeignPtr Double
cplxPtr = castForeignPtr . toForeignPtrs

newCplx :: Int -> IO (ForeignPtr Double)
newCplx n = mallocForeignPtrArray (2 * n)

-- | frac_coords (src/Gridding.hs:142-151): (height, width) qpx (u, v) -> (x, xf, y, yf)
fracCoords :: Ctx -> (Int, Int) -> Int -> Vector F -> Vector F -> IO (Vector Int64, Vector Int64, Vector Int64, Vector Int64)
fracCoords ctx (h, w) qpx u v = do
  let n = arraySize u
  [x, xf, y, yf] <- mapM (const (mallocForeignPtrArray n)) [1 :: Int .. 4]
  withForeignPtr (toForeignPtrs u) $ \pu -> withForeignPtr (toForeignPtrs v) $ \pv ->
    withForeignPtr x $ \px -> withForeignPtr xf $ \pxf -> withForeignPtr y $ \py -> withForeignPtr yf $ \pyf ->
      c_frac_coords ctx (fromIntegral h) (fromIntegral w) (fromIntegral qpx) (fromIntegral n) pu pv px pxf py pyf 1
        >>= check ctx "frac_coords"
  let mk = fromForeignPtrs (Z :. n)
  return (mk x, mk xf, mk y, mk yf)

-- | findClosest (src/Gridding.hs:895-907) over a vector of w
findClosestV :: Ctx -> Vector F -> Vector F -> IO (Vector Int64)
findClosestV ctx ws w = do
  let n = arraySize w
  out <- mallocForeignPtrArray n
  withForeignPtr (toForeignPtrs ws) $ \pws -> withForeignPtr (toForeignPtrs w) $ \pw -> withForeignPtr out $ \po ->
    c_find_closest ctx (fromIntegral (arraySize ws)) pws (fromIntegral n) pw po >>= check ctx "findClosest"
  return (fromForeignPtrs (Z :. n) out)

-- | mirror_uvw (src/Gridding.hs:551-562), in place on copies made by the caller
mirrorUvw :: Ctx -> Int -> ForeignPtr Double -> ForeignPtr Double -> ForeignPtr Double -> ForeignPtr Double -> IO ()
mirrorUvw ctx n u v w vis =
  withForeignPtr u $ \pu -> withForeignPtr v $ \pv -> withForeignPtr w $ \pw -> withForeignPtr vis $ \pvis ->
    c_mirror_uvw ctx (fromIntegral n) pu pv pw pvis >>= check ctx "mirror_uvw"

-- | doweight (src/Gridding.hs:564-583), vis divided in place
doweight :: Ctx -> F -> Int -> Vector F -> Vector F -> ForeignPtr Double -> IO ()
doweight ctx theta lam u v vis =
  withForeignPtr (toForeignPtrs u) $ \pu -> withForeignPtr (toForeignPtrs v) $ \pv -> withForeignPtr vis $ \pvis ->
    c_doweight ctx (CDouble theta) (fromIntegral lam) (fromIntegral (arraySize u)) pu pv pvis >>= check ctx "doweight"

-- | convgrid (src/Gridding.hs:153-197): gcf [qpx,qpx,gh,gw], grid accumulated in place
convgrid :: Ctx -> Array (Z :. Int :. Int :. Int :. Int) Visibility -> (Int, Int) -> ForeignPtr Double
         -> Vector F -> Vector F -> Vector Visibility -> IO ()
convgrid ctx gcf (h, w) grid u v vis = do
  let Z :. qpx :. _ :. gh :. gw = arrayShape gcf
  withForeignPtr (cplxPtr gcf) $ \pk -> withForeignPtr grid $ \pg -> withForeignPtr (toForeignPtrs u) $ \pu ->
    withForeignPtr (toForeignPtrs v) $ \pv -> withForeignPtr (cplxPtr vis) $ \pvis ->
      c_convgrid ctx (fromIntegral qpx) (fromIntegral gh) (fromIntegral gw) pk (fromIntegral h) (fromIntegral w) pg
                 (fromIntegral (arraySize u)) pu pv pvis >>= check ctx "convgrid"

-- | convgrid2 (src/Gridding.hs:199-244): gcf [nw,qpx,qpx,gh,gw] + wbin
convgrid2 :: Ctx -> Array DIM5 Visibility -> (Int, Int) -> ForeignPtr Double
          -> Vector F -> Vector F -> Vector Int64 -> Vector Visibility -> IO ()
convgrid2 ctx gcf (h, w) grid u v wbin vis = do
  let Z :. nw :. qpx :. _ :. gh :. gw = arrayShape gcf
  withForeignPtr (cplxPtr gcf) $ \pk -> withForeignPtr grid $ \pg -> withForeignPtr (toForeignPtrs u) $ \pu ->
    withForeignPtr (toForeignPtrs v) $ \pv -> withForeignPtr (toForeignPtrs wbin) $ \pwb -> withForeignPtr (cplxPtr vis) $ \pvis ->
      c_convgrid2 ctx (fromIntegral nw) (fromIntegral qpx) (fromIntegral gh) (fromIntegral gw) pk (fromIntegral h) (fromIntegral w) pg
                  (fromIntegral (arraySize u)) pu pv pwb pvis >>= check ctx "convgrid2"

-- | convgrid3 / convgrid4 (src/Gridding.hs:246-396): wkerns [nw,qpx,qpx,s,s], akerns [nant,s,s], index (wbin,a1,a2)
convgridAW :: Ctx -> Array DIM5 Visibility -> Array DIM3 Visibility -> (Int, Int) -> ForeignPtr Double
           -> Vector F -> Vector F -> (Vector Int64, Vector Int64, Vector Int64) -> Vector Visibility -> IO ()
convgridAW ctx wk ak (h, w) grid u v (wbin, a1, a2) vis = do
  let Z :. nw :. qpx :. _ :. s :. _ = arrayShape wk
      Z :. nant :. _ :. _ = arrayShape ak
  withForeignPtr (cplxPtr wk) $ \pwk -> withForeignPtr (cplxPtr ak) $ \pak -> withForeignPtr grid $ \pg ->
    withForeignPtr (toForeignPtrs u) $ \pu -> withForeignPtr (toForeignPtrs v) $ \pv ->
    withForeignPtr (toForeignPtrs wbin) $ \pwb -> withForeignPtr (toForeignPtrs a1) $ \pa1 -> withForeignPtr (toForeignPtrs a2) $ \pa2 ->
    withForeignPtr (cplxPtr vis) $ \pvis ->
      c_convgrid_aw ctx (fromIntegral nw) (fromIntegral qpx) (fromIntegral s) pwk (fromIntegral nant) pak (fromIntegral h) (fromIntegral w) pg
                    (fromIntegral (arraySize u)) pu pv pwb pa1 pa2 pvis >>= check ctx "convgrid_aw"

-- | adjoint of convgrid2 (new: the reference has no degridding)
convdegrid2 :: Ctx -> Array DIM5 Visibility -> Array DIM2 Visibility -> Vector F -> Vector F -> Vector Int64 -> IO (Vector Visibility)
convdegrid2 ctx gcf grid u v wbin = do
  let Z :. nw :. qpx :. _ :. gh :. gw = arrayShape gcf
      Z :. h :. w = arrayShape grid
      n = arraySize u
  out <- newCplx n
  withForeignPtr (cplxPtr gcf) $ \pk -> withForeignPtr (cplxPtr grid) $ \pg -> withForeignPtr (toForeignPtrs u) $ \pu ->
    withForeignPtr (toForeignPtrs v) $ \pv -> withForeignPtr (toForeignPtrs wbin) $ \pwb -> withForeignPtr out $ \po ->
      c_convdegrid2 ctx (fromIntegral nw) (fromIntegral qpx) (fromIntegral gh) (fromIntegral gw) pk (fromIntegral h) (fromIntegral w) pg
                    (fromIntegral n) pu pv pwb po >>= check ctx "convdegrid2"
  return (fromForeignPtrs (Z :. n) (castForeignPtr out))

-- | One context per CUDA device, for the multi-GPU calls below (errors are reported through the first context).
withSkaGrids :: [Int] -> ([Ctx] -> IO a) -> IO a
withSkaGrids []       k = k []
withSkaGrids (d : ds) k = withSkaGrid d $ \c -> withSkaGrids ds (k . (c :))

-- | VisSharded: contiguous shares of the visibilities + reduce (BASELINE config 4);
--   TileSharded: row slabs of the grid + device-to-device routing (config 5).
data MultiMode = VisSharded | TileSharded

-- | convgrid2 over several devices, driven by this one thread; the grid is accumulated in place.
convgrid2Multi :: MultiMode -> [Ctx] -> Array DIM5 Visibility -> (Int, Int) -> ForeignPtr Double
               -> Vector F -> Vector F -> Vector Int64 -> Vector Visibility -> IO ()
convgrid2Multi mode ctxs gcf (h, w) grid u v wbin vis = do
  let Z :. nw :. qpx :. _ :. gh :. gw = arrayShape gcf
      i = fromIntegral
  withArrayLen ctxs $ \nc pc -> withForeignPtr (cplxPtr gcf) $ \pk -> withForeignPtr grid $ \pg ->
    withForeignPtr (toForeignPtrs u) $ \pu -> withForeignPtr (toForeignPtrs v) $ \pv ->
    withForeignPtr (toForeignPtrs wbin) $ \pwb -> withForeignPtr (cplxPtr vis) $ \pvis ->
      (case mode of
         VisSharded  -> c_convgrid2_mgpu_vis  pc (i nc) (i nw) (i qpx) (i gh) (i gw) pk (i h) (i w) pg (i (arraySize u)) pu pv pwb pvis
         TileSharded -> c_convgrid2_mgpu_tile pc (i nc) (i nw) (i qpx) (i gh) (i gw) pk (i h) (i w) pg (i (arraySize u)) pu pv pwb pvis nullPtr)
        >>= check (head ctxs) "convgrid2Multi"

-- | adjoint of convgrid2Multi
convdegrid2Multi :: MultiMode -> [Ctx] -> Array DIM5 Visibility -> Array DIM2 Visibility
                 -> Vector F -> Vector F -> Vector Int64 -> IO (Vector Visibility)
convdegrid2Multi mode ctxs gcf grid u v wbin = do
  let Z :. nw :. qpx :. _ :. gh :. gw = arrayShape gcf
      Z :. h :. w = arrayShape grid
      n = arraySize u
      i = fromIntegral
  out <- newCplx n
  withArrayLen ctxs $ \nc pc -> withForeignPtr (cplxPtr gcf) $ \pk -> withForeignPtr (cplxPtr grid) $ \pg ->
    withForeignPtr (toForeignPtrs u) $ \pu -> withForeignPtr (toForeignPtrs v) $ \pv ->
    withForeignPtr (toForeignPtrs wbin) $ \pwb -> withForeignPtr out $ \po ->
      (case mode of
         VisSharded  -> c_convdegrid2_mgpu_vis  pc (i nc) (i nw) (i qpx) (i gh) (i gw) pk (i h) (i w) pg (i n) pu pv pwb po
         TileSharded -> c_convdegrid2_mgpu_tile pc (i nc) (i nw) (i qpx) (i gh) (i gw) pk (i h) (i w) pg (i n) pu pv pwb po nullPtr)
        >>= check (head ctxs) "convdegrid2Multi"
  return (fromForeignPtrs (Z :. n) (castForeignPtr out))

-- | adjoint of convgrid3/4
convdegridAW :: Ctx -> Array DIM5 Visibility -> Array DIM3 Visibility -> Array DIM2 Visibility
             -> Vector F -> Vector F -> (Vector Int64, Vector Int64, Vector Int64) -> IO (Vector Visibility)
convdegridAW ctx wk ak grid u v (wbin, a1, a2) = do
  let Z :. nw :. qpx :. _ :. s :. _ = arrayShape wk
      Z :. nant :. _ :. _ = arrayShape ak
      Z :. h :. w = arrayShape grid
      n = arraySize u
  out <- newCplx n
  withForeignPtr (cplxPtr wk) $ \pwk -> withForeignPtr (cplxPtr ak) $ \pak -> withForeignPtr (cplxPtr grid) $ \pg ->
    withForeignPtr (toForeignPtrs u) $ \pu -> withForeignPtr (toForeignPtrs v) $ \pv ->
    withForeignPtr (toForeignPtrs wbin) $ \pwb -> withForeignPtr (toForeignPtrs a1) $ \pa1 -> withForeignPtr (toForeignPtrs a2) $ \pa2 ->
    withForeignPtr out $ \po ->
      c_convdegrid_aw ctx (fromIntegral nw) (fromIntegral qpx) (fromIntegral s) pwk (fromIntegral nant) pak (fromIntegral h) (fromIntegral w) pg
                      (fromIntegral n) pu pv pwb pa1 pa2 po >>= check ctx "convdegrid_aw"
  return (fromForeignPtrs (Z :. n) (castForeignPtr out))

-- | make_grid_hermitian (src/Gridding.hs:585-605)
makeGridHermitian :: Ctx -> Array DIM2 Visibility -> IO (Array DIM2 Visibility)
makeGridHermitian ctx g = do
  let sh@(Z :. n :. _) = arrayShape g
  out <- newCplx (n * n)
  withForeignPtr (cplxPtr g) $ \pg -> withForeignPtr out $ \po -> c_hermitian ctx (fromIntegral n) pg po >>= check ctx "make_grid_hermitian"
  return (fromForeignPtrs sh (castForeignPtr out))

-- | ifft (src/Gridding.hs:828-829)
ifft :: Ctx -> Array DIM2 Visibility -> IO (Array DIM2 Visibility)
ifft ctx g = do
  let sh@(Z :. n :. _) = arrayShape g
  out <- newCplx (n * n)
  withForeignPtr (cplxPtr g) $ \pg -> withForeignPtr out $ \po -> c_ifft ctx (fromIntegral n) pg po >>= check ctx "ifft"
  return (fromForeignPtrs sh (castForeignPtr out))

-- | map real . ifft . make_grid_hermitian, and its maximum (src/ImageDataset.hs:74-77)
gridToImage :: Ctx -> Array DIM2 Visibility -> IO (Array DIM2 F, F)
gridToImage ctx g = do
  let sh@(Z :. n :. _) = arrayShape g
  img <- mallocForeignPtrArray (n * n)
  mx <- alloca $ \pm -> do
    withForeignPtr (cplxPtr g) $ \pg -> withForeignPtr img $ \pi' -> c_grid_to_image ctx (fromIntegral n) pg pi' pm >>= check ctx "grid_to_image"
    peek pm
  return (fromForeignPtrs sh img, mx)

-- | aw_imaging (src/Gridding.hs:452-478): u,v,w in wavelengths -> uv grid
awImaging :: Ctx -> F -> Int -> Array DIM5 Visibility -> Vector F -> Array DIM3 Visibility
          -> (Vector F, Vector F, Vector F) -> (Vector Int64, Vector Int64) -> Vector Visibility -> IO (Array DIM2 Visibility)
awImaging ctx theta lam wk wbins ak (u, v, w) (a1, a2) vis = do
  let Z :. nw :. qpx :. _ :. s :. _ = arrayShape wk
      Z :. nant :. _ :. _ = arrayShape ak
      n = round (theta * fromIntegral lam) :: Int
  out <- newCplx (n * n)
  withForeignPtr (cplxPtr wk) $ \pwk -> withForeignPtr (toForeignPtrs wbins) $ \pwb -> withForeignPtr (cplxPtr ak) $ \pak ->
    withForeignPtr (toForeignPtrs u) $ \pu -> withForeignPtr (toForeignPtrs v) $ \pv -> withForeignPtr (toForeignPtrs w) $ \pw ->
    withForeignPtr (toForeignPtrs a1) $ \pa1 -> withForeignPtr (toForeignPtrs a2) $ \pa2 -> withForeignPtr (cplxPtr vis) $ \pvis ->
    withForeignPtr out $ \po ->
      c_aw_imaging ctx (CDouble theta) (fromIntegral lam) (fromIntegral nw) (fromIntegral qpx) (fromIntegral s) pwk pwb (fromIntegral nant) pak
                   (fromIntegral (arraySize u)) pu pv pw pa1 pa2 pvis po >>= check ctx "aw_imaging"
  return (fromForeignPtrs (Z :. n :. n) (castForeignPtr out))

-- | ImageDataset.aw_gridding from the loaded arrays on (src/ImageDataset.hs:47-77): uvw in metres; returns the image
-- and its maximum (what app/Main.hs:54 prints).
awGridding :: Ctx -> F -> Int -> Array DIM5 Visibility -> Vector F -> Array DIM3 Visibility
           -> (Vector F, Vector F, Vector F) -> (Vector Int64, Vector Int64) -> F -> Vector Visibility -> IO (Array DIM2 F, F)
awGridding ctx theta lam wk wbins ak (u, v, w) (a1, a2) freq vis = do
  let Z :. nw :. qpx :. _ :. s :. _ = arrayShape wk
      Z :. nant :. _ :. _ = arrayShape ak
      n = round (theta * fromIntegral lam) :: Int
  img <- mallocForeignPtrArray (n * n)
  mx <- alloca $ \pm -> do
    withForeignPtr (cplxPtr wk) $ \pwk -> withForeignPtr (toForeignPtrs wbins) $ \pwb -> withForeignPtr (cplxPtr ak) $ \pak ->
      withForeignPtr (toForeignPtrs u) $ \pu -> withForeignPtr (toForeignPtrs v) $ \pv -> withForeignPtr (toForeignPtrs w) $ \pw ->
      withForeignPtr (toForeignPtrs a1) $ \pa1 -> withForeignPtr (toForeignPtrs a2) $ \pa2 -> withForeignPtr (cplxPtr vis) $ \pvis ->
      withForeignPtr img $ \pi' ->
        c_aw_gridding ctx (CDouble theta) (fromIntegral lam) (fromIntegral nw) (fromIntegral qpx) (fromIntegral s) pwk pwb (fromIntegral nant) pak
                      (fromIntegral (arraySize u)) pu pv pw pa1 pa2 (CDouble freq) pvis pi' pm nullPtr >>= check ctx "aw_gridding"
    peek pm
  return (fromForeignPtrs (Z :. n :. n) img, mx)
